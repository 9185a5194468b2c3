"""GPU: the optimize_acqf stand-in as a GPU workload (reference call sites
mobocmf/acquisition_functions/JESMOC_MFDGP.py:142-143,151-168): restarts x fidelities in one batch, value and gradient
replayed from a CUDA graph."""
import numpy as np
import pytest
import torch

from tests.helpers import model_from_state, random_state

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _pair(M, d, L, seed, ls):
    sd, up = random_state(M, d, L, seed=seed, ls=ls)
    sdc = {k: v.clone() for k, v in sd.items()}
    g = torch.Generator().manual_seed(seed + 1)
    for k in sdc:
        if "chol_variational_covar" in k:
            sdc[k] = sdc[k] * 0.5
        if "variational_mean" in k:
            sdc[k] = sdc[k] + 0.05 * torch.randn(sdc[k].shape, generator=g, dtype=torch.float64)
    samples = [torch.randn(7, 1, generator=g) for _ in range(L)]
    return model_from_state(sd, up, L, samples=samples), model_from_state(sdc, up, L, samples=samples)


def _acqs(L, d, seed=3, M=20, ls=0.25):
    from mobocmf_b200.acquisition_functions.JESMOC_MFDGP import _JES_MFDGP
    u, c = _pair(M, d, L, seed, ls)
    u2, c2 = _pair(M, d, L, seed + 10, ls)
    fns = []
    for f in range(L):
        j1, j2 = _JES_MFDGP(f, u, c), _JES_MFDGP(f, u2, c2)
        fns.append(lambda X, j1=j1, j2=j2: j1(X) + j2(X))        # two black boxes, like coupled_acq's sum
    return fns


def test_graph_replay_equals_eager_value_and_gradient():
    from mobocmf_b200.util.optimize import GraphedValueAndGrad
    L, d, nb = 2, 2, 5
    fns = _acqs(L, d)

    def joint(X):
        return torch.cat([fn(X[f * nb:(f + 1) * nb]).double() for f, fn in enumerate(fns)])
    g = GraphedValueAndGrad(joint, L * nb, d, torch.device(DEV))
    for seed in (0, 1):
        x = torch.rand(L * nb, 1, d, generator=torch.Generator().manual_seed(seed), dtype=torch.float64)
        loss, grad, vals = g(x.numpy().reshape(-1))
        X = x.to(DEV).requires_grad_(True)
        v = joint(X)
        le = -v.sum()
        ge, = torch.autograd.grad(le, X)
        assert abs(loss - float(le)) <= 1e-12 * max(1.0, abs(float(le)))
        assert np.abs(grad - ge.cpu().numpy().reshape(-1)).max() <= 1e-12 * max(1.0, float(ge.abs().max()))
        assert np.abs(vals - v.detach().cpu().numpy()).max() <= 1e-12
        assert float(ge.abs().max()) > 0.0


def test_multi_fidelity_optimizer_finds_the_grid_maximum():
    """1-D inputs: a dense grid is the known answer.  Every fidelity's optimum, found in ONE joint L-BFGS-B run from a
    CUDA graph, must reach the best grid value (the grid has spacing 2.5e-4, the optimiser refines beyond it)."""
    from mobocmf_b200.util.optimize import optimize_acqf_multi
    L, d = 2, 1
    fns = _acqs(L, d, seed=5, M=16, ls=0.15)
    bounds = torch.tensor([[0.0], [1.0]], dtype=torch.float64, device=DEV)
    out, info = optimize_acqf_multi(fns, bounds, num_restarts=5, raw_samples=128, options={"maxiter": 100}, seed=0,
                                    return_info=True)
    assert info["graph"] and info["evaluations"] >= 3
    grid = torch.linspace(0, 1, 4001, dtype=torch.float64, device=DEV)[:, None, None]
    for f, (cand, val) in enumerate(out):
        with torch.no_grad():
            gv = fns[f](grid).double()
            again = fns[f](cand.reshape(1, 1, d)).double()
        assert cand.shape == (1, d) and 0.0 <= float(cand) <= 1.0
        assert abs(float(again) - float(val)) <= 1e-12 * max(1.0, abs(float(val)))
        assert float(val) >= float(gv.max()) - 1e-9, (f, float(val), float(gv.max()))
        assert float(gv.max()) > 1e-3          # a non-trivial acquisition surface
    # the eager path (no graph) lands on the same optima
    out_e = optimize_acqf_multi(fns, bounds, num_restarts=5, raw_samples=128, options={"maxiter": 100}, seed=0,
                                use_cuda_graph=False)
    for (c1, v1), (c2, v2) in zip(out, out_e):
        assert abs(float(v1) - float(v2)) <= 1e-8 * max(1.0, abs(float(v1)))
