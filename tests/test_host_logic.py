"""CPU tests of the host-side mirror of the reference interface (no kernels are launched): GPyTorch-compatible
parameter naming, constraint transforms, deep-copy / pickle safety of the objects the fitter copies
(mobocmf/util/blackbox_mfdgp_fitter.py:372-397), the optimize_acqf stand-in, the operator-buffer layout mirror."""
import copy
import pickle

import torch

from tests.helpers import forrester_data


def _model(L=2):
    from mobocmf_b200.models.mfdgp import MFDGP
    x, ys, fid = forrester_data()
    torch.manual_seed(0)
    m = MFDGP(x, ys["obj1"], fid, L)
    m.double()
    return m, x, ys, fid


def test_state_dict_uses_gpytorch_names():
    m, *_ = _model()
    keys = set(m.state_dict().keys())
    expect = {
        "hidden_layer_0.variational_strategy.inducing_points",
        "hidden_layer_0.variational_strategy.variational_params_initialized",
        "hidden_layer_0.variational_strategy._variational_distribution.variational_mean",
        "hidden_layer_0.variational_strategy._variational_distribution.chol_variational_covar",
        "hidden_layer_0.covar_module.raw_outputscale",
        "hidden_layer_0.covar_module.base_kernel.raw_lengthscale",
        "hidden_layer_1.covar_module.kernels.0.kernels.0.raw_outputscale",
        "hidden_layer_1.covar_module.kernels.0.kernels.0.base_kernel.raw_lengthscale",
        "hidden_layer_1.covar_module.kernels.0.kernels.1.kernels.0.raw_variance",
        "hidden_layer_1.covar_module.kernels.0.kernels.1.kernels.1.raw_outputscale",
        "hidden_layer_1.covar_module.kernels.0.kernels.1.kernels.1.base_kernel.raw_lengthscale",
        "hidden_layer_1.covar_module.kernels.1.raw_outputscale",
        "hidden_layer_1.covar_module.kernels.1.base_kernel.raw_lengthscale",
        "hidden_layer_likelihood_0.noise_covar.raw_noise",
        "hidden_layer_likelihood_1.noise_covar.raw_noise",
        # quirk Q12: the previous layer is registered inside the strategy of the next one
        "hidden_layer_1.variational_strategy.previous_layer.covar_module.raw_outputscale",
    }
    missing = expect - keys
    assert not missing, missing
    sd = m.state_dict()
    assert sd["hidden_layer_0.covar_module.base_kernel.raw_lengthscale"].shape == (1, 1)
    assert sd["hidden_layer_1.covar_module.kernels.0.kernels.1.kernels.0.raw_variance"].shape == (1, 1)
    assert sd["hidden_layer_1.variational_strategy.inducing_points"].shape == (16, 2)
    assert sd["hidden_layer_likelihood_1.noise_covar.raw_noise"].shape == (1,)
    # parameters() de-duplicates the shared previous layer (KL counts each layer once)
    assert len(list(m.parameters())) == 2 + 7 + 2 * 2 + 2
    m2, *_ = _model()
    m2.load_state_dict(sd)


def test_freeze_helpers_match_reference_semantics():
    m, *_ = _model()
    m.fix_variational_hypers(True)      # phase 1: noise and chol_variational_covar frozen (models/mfdgp.py:198-206)
    frozen = {n for n, p in m.named_parameters() if not p.requires_grad}
    assert all(("raw_noise" in n) or ("chol_variational_covar" in n) for n in frozen) and len(frozen) == 4
    m.fix_variational_hypers(False)
    m.fix_variational_hypers_cond(True)  # conditioned training: noise + every kernel hyper-parameter frozen (:208-218)
    free = {n for n, p in m.named_parameters() if p.requires_grad}
    assert all("_variational_distribution" in n for n in free) and len(free) == 4


def test_constraint_transforms_round_trip():
    from mobocmf_b200.gp import Interval, Positive
    it = Interval(1e-8, 0.3)
    v = torch.tensor([1e-6, 0.01, 0.25], dtype=torch.float64)
    assert torch.allclose(it.transform(it.inverse_transform(v)), v, rtol=1e-6)      # float32 bounds like upstream
    pos = Positive()
    assert torch.allclose(pos.transform(pos.inverse_transform(v)), v, rtol=1e-12)
    m, *_ = _model()
    # initial noise values (models/mfdgp.py:118-121): 1e-6 on the lower layer
    assert abs(float(m.hidden_layer_likelihood_0.noise) - 1e-6) < 1e-9


def test_elbo_and_model_survive_deepcopy_and_pickle_with_live_bindings():
    from mobocmf_b200.mlls.variational_elbo_mf import VariationalELBOMF
    m, x, ys, fid = _model()
    elbo = VariationalELBOMF(m, 16, 2)

    class NotPicklable(object):
        def __reduce_ex__(self, protocol):
            raise ValueError("ctypes objects containing pointers cannot be pickled")
    elbo._fused_step = NotPicklable()                  # what the fitter caches after the first fused step
    m.hidden_layer_0._ops_cache = ("key", NotPicklable())
    e2 = copy.deepcopy(elbo)
    assert not hasattr(e2, "_fused_step") and e2.num_data == 16
    m2 = copy.deepcopy(m)
    assert m2.hidden_layer_0._ops_cache is None
    m3 = pickle.loads(pickle.dumps(m))
    assert torch.equal(m3.hidden_layer_1.samples, m.hidden_layer_1.samples)       # common random numbers (quirk Q7)


def test_optimize_acqf_stand_in_finds_the_maximum_on_cpu():
    from mobocmf_b200.util.optimize import optimize_acqf
    target = torch.tensor([0.3, 0.7], dtype=torch.double)

    def acq(X):                      # (b, 1, d) -> (b,), BoTorch's contract (acquisition_functions/...py:142-143)
        assert X.dim() == 3 and X.shape[1] == 1
        return -((X[:, 0, :] - target) ** 2).sum(-1)
    bounds = torch.tensor([[0.0, 0.0], [1.0, 1.0]], dtype=torch.double)
    x, v = optimize_acqf(acq, bounds, q=1, num_restarts=5, raw_samples=64, options={"maxiter": 50}, seed=0)
    assert x.shape == (1, 2) and torch.allclose(x[0], target, atol=1e-5) and float(v) > -1e-9


def test_optimize_acqf_multi_equals_separate_runs_on_cpu():
    """get_nextpoint_coupled's loop over the fidelities (acquisition_functions/JESMOC_MFDGP.py:151-168) as ONE
    multi-start L-BFGS-B run: the joint objective is separable, so every function ends at its own maximum."""
    from mobocmf_b200.util.optimize import optimize_acqf, optimize_acqf_multi
    targets = [torch.tensor([0.3, 0.7], dtype=torch.double), torch.tensor([0.8, 0.1], dtype=torch.double),
               torch.tensor([0.5, 0.5], dtype=torch.double)]
    calls = []

    def make(t, scale):
        def acq(X):
            assert X.dim() == 3 and X.shape[1] == 1
            calls.append(X.shape[0])
            return scale - scale * ((X[:, 0, :] - t) ** 2).sum(-1)
        return acq
    fns = [make(t, s) for t, s in zip(targets, (1.0, 3.0, 0.5))]
    bounds = torch.tensor([[0.0, 0.0], [1.0, 1.0]], dtype=torch.double)
    out, info = optimize_acqf_multi(fns, bounds, num_restarts=4, raw_samples=32, options={"maxiter": 60}, seed=1,
                                    return_info=True)
    assert len(out) == 3 and info["graph"] is False and info["evaluations"] >= 2
    for (x, v), t, s in zip(out, targets, (1.0, 3.0, 0.5)):
        assert x.shape == (1, 2) and torch.allclose(x[0], t, atol=1e-5) and abs(float(v) - s) < 1e-8
    assert set(calls) == {32, 4}          # raw-sample screening, then the restarts of each function as one slice
    x1, v1 = optimize_acqf(fns[1], bounds, num_restarts=4, raw_samples=32, options={"maxiter": 60}, seed=1)
    assert torch.allclose(x1[0], out[1][0][0], atol=1e-5)


def test_operator_buffer_layout_mirror_matches_library():
    from mobocmf_b200 import _lib
    from mobocmf_b200.functional import ops_layout, padded_m
    lib = _lib.load()
    for M in (1, 16, 32, 33, 75, 256):
        lay = ops_layout(M)
        assert lay["MP"] == lib.mobo_padded_m(M) == padded_m(M)
        assert lay["size"] == lib.mobo_ops_doubles(M)


def test_small_host_helpers(tmp_path):
    """Host helpers mirrored from mobocmf/util/util.py and models/mfdgp.py:125-135."""
    import numpy as np
    from mobocmf_b200.models.mfdgp import MFDGP
    from mobocmf_b200.util import util
    a, b, mean, std = util.preprocess_outputs(np.array([[1.0], [2.0]]), np.array([[3.0]]))
    assert a.dtype == torch.float64 and float(a[1]) == 2.0 and float(b[0]) == 3.0 and (mean, std) == (0.0, 1.0)
    lo, hi, mean, std = util.preprocess_outputs_two_fidelities(np.array([[1.0]]), np.array([[4.0]]))
    assert float(lo) == 1.0 and float(hi) == 4.0
    util.save_pickle(str(tmp_path / "sub"), "x.pkl", {"k": torch.arange(3)})
    assert torch.equal(util.read_pickle(str(tmp_path / "sub"), "x.pkl")["k"], torch.arange(3))
    x0 = torch.tensor([[0.0], [0.9], [0.52]], dtype=torch.float64)
    x1 = torch.tensor([[0.1], [0.5], [1.0]], dtype=torch.float64)
    y1 = torch.tensor([[10.0], [20.0], [30.0]], dtype=torch.float64)
    assert torch.equal(MFDGP.clip_inducing_values(None, x0, x1, y1), torch.tensor([[10.0], [30.0], [20.0]],
                                                                                dtype=torch.float64))
