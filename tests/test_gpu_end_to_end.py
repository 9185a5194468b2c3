"""End-to-end flow of the reference's Forrester example (examples/example_acquisition_mfdgp_forrester/...py, config C2)
through the drop-in classes on the GPU: fit 3 black boxes (2 objectives + 1 constraint), conditioned training on a
given Pareto set, coupled acquisition on a grid, next-point selection.  Orchestration only (the arithmetic is covered
by the parity tests); checks shapes, finiteness, that training lowers the loss and that both step paths (eager fused,
CUDA graph) are exercised by the fitter."""
import pytest
import torch

from tests.helpers import forrester_data

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("use_cuda_graph,concurrent", [(False, False), (True, False), (True, True)])
def test_forrester_fit_condition_acquire(use_cuda_graph, concurrent):
    from mobocmf_b200.acquisition_functions.JESMOC_MFDGP import JESMOC_MFDGP
    from mobocmf_b200.util.blackbox_mfdgp_fitter import BlackBoxMFDGPFitter
    x, ys, fid = forrester_data()
    L, N = 2, x.shape[0]
    torch.manual_seed(0)
    fitter = BlackBoxMFDGPFitter(L, N, num_epochs_1=40, num_epochs_2=40, device=torch.device(DEV),
                                 use_cuda_graph=use_cuda_graph, concurrent_models=concurrent, pareto_set_size=6,
                                 opt_grid_size=150)
    fitter.verbose = False
    fitter.initialize_mfdgp(x, ys["obj1"], fid, "obj1")
    fitter.initialize_mfdgp(x, ys["obj2"], fid, "obj2")
    fitter.initialize_mfdgp(x, ys["con1"], fid, "con1", threshold_constraint=0.0, is_constraint=True)
    h = fitter.mfdgp_handlers_objs["obj1"]
    from mobocmf_b200.gp import settings
    with torch.no_grad(), settings.num_likelihood_samples(1):
        before = float(-h.elbo(h.mfdgp(h.x, eps=[None, torch.zeros(N, device=DEV)]), h.y.T, h.f)[0])
    fitter.train_mfdgps()
    with torch.no_grad(), settings.num_likelihood_samples(1):
        after = float(-h.elbo(h.mfdgp(h.x, eps=[None, torch.zeros(N, device=DEV)]), h.y.T, h.f)[0])
    assert after < before, (before, after)
    if use_cuda_graph:
        assert any(getattr(hh.elbo, "_fused_step", None) and getattr(hh.elbo._fused_step, "_graphs", None)
                   for hh in fitter.mfdgp_handlers_objs.values())
    if concurrent:
        # the three models trained on their own streams: every one of them moved and stayed finite
        for hh in list(fitter.mfdgp_handlers_objs.values()) + list(fitter.mfdgp_handlers_cons.values()):
            assert all(bool(torch.isfinite(p).all()) for p in hh.mfdgp.parameters())
        # the Pareto set is drawn by JESMOC_MFDGP itself below, like in the reference
        # (acquisition_functions/JESMOC_MFDGP.py:73-77): RFF function samples + MOOP on the GPU
        import numpy as np
        np.random.seed(4)
    else:
        g = torch.Generator().manual_seed(1)
        fitter.pareto_set = torch.rand(6, 1, generator=g, dtype=torch.float64)
        fitter.pareto_front = torch.randn(6, 2, generator=g, dtype=torch.float64) * 0.3
    fitter.num_epochs_2 = 15
    bounds = torch.tensor([[0.0], [1.0]], dtype=torch.float64, device=DEV)
    acq = JESMOC_MFDGP(model=fitter, num_fidelities=L, standard_bounds=bounds)
    if concurrent:
        pset, pfront = acq.pareto_set, acq.pareto_front
        assert pset.shape[1] == 1 and pfront.shape == (pset.shape[0], 2) and 1 <= pset.shape[0] <= 6
    for f in range(L):
        acq.add_blackbox(f, "obj1", cost_evaluation=1.0 + f)
        acq.add_blackbox(f, "obj2", cost_evaluation=1.0 + f)
        acq.add_blackbox(f, "con1", cost_evaluation=1.0 + f, is_constraint=True)
    grid = torch.linspace(0, 1, 200, dtype=torch.float64, device=DEV)[:, None, None]
    for f in range(L):
        with torch.no_grad():
            v = acq.coupled_acq(grid, f)
        assert v.shape == (200,) and bool(torch.isfinite(v).all()) and float(v.min()) >= 0.0
        assert v.dtype == torch.float32                     # quirk Q8: the reference accumulates in float32
        d = acq.decoupled_acq(grid, f, "con1", is_constraint=True)
        assert d.shape == (200,) and d.dtype == torch.float64
    nextpoint, fidelity = acq.get_nextpoint_coupled(iteration=0, verbose=False)
    assert nextpoint.shape == (1,) and 0.0 <= float(nextpoint) <= 1.0 and fidelity in (0, 1)
