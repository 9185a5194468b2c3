"""CPU: model construction of the product (mobocmf_b200.models.mfdgp.MFDGP) against the oracle's restatement of
models/mfdgp.py:22-151,290-317 and layers/mfdgp_hidden_layer.py:26-161, on the reference's Forrester fixture."""
import numpy as np
import torch

from oracle import mfdgp_oracle as O
from tests.helpers import forrester_data, synthetic_data, oracle_view


def test_forrester_known_lengthscales():
    # TL.MEDIAN with quirk Q2 gives 0.2727... for the LF inputs and 0.2 for the HF inputs
    x, ys, fid = forrester_data()
    assert abs(float(O.init_lengthscale_median(x[fid.flatten() == 0])) - 3.0 / 11.0) < 1e-12
    assert abs(float(O.init_lengthscale_median(x[fid.flatten() == 1])) - 0.2) < 1e-12


def _check_init(x, y, fid, L):
    from mobocmf_b200.models.mfdgp import MFDGP
    torch.manual_seed(3)
    model = MFDGP(x, y, fid, L)
    model.double()
    torch.manual_seed(3)
    sd_o, lo_o, up_o, _ = O.init_state_dict(x, y, fid, L)
    sd_p, lo_p, up_p, samples = oracle_view(model)
    for k, v in sd_o.items():
        assert k in sd_p, k
        assert sd_p[k].shape == v.shape, k
        if "chol_variational_covar" in k:
            # float32-rounded Cholesky of a (possibly ill-conditioned) K: compare the covariance it encodes
            a, b = sd_p[k] @ sd_p[k].T, v @ v.T
            assert (a - b).abs().max() <= 1e-5 * b.abs().max(), k
        elif v.dtype.is_floating_point:
            assert torch.allclose(sd_p[k], v, rtol=1e-6, atol=1e-12), k
        else:
            assert torch.equal(sd_p[k], v), k
    assert np.allclose(lo_o, lo_p) and np.allclose(up_o, up_p)
    assert all(s.shape == (25, 1) and s.dtype == torch.float32 for s in samples)
    # quirk Q1: the initial variational means are float32 roundings
    for l in range(L):
        m = sd_p["hidden_layer_%d.variational_strategy._variational_distribution.variational_mean" % l]
        assert torch.equal(m, m.float().double())


def test_init_matches_oracle_forrester():
    x, ys, fid = forrester_data()
    for name in ("obj1", "obj2", "con1"):
        _check_init(x, ys[name], fid, 2)


def test_init_matches_oracle_three_fidelities():
    x, y, fid = synthetic_data([20, 12, 8], 3, seed=1)
    _check_init(x, y, fid, 3)


def test_state_dict_has_gpytorch_names_and_duplicates():
    from mobocmf_b200.models.mfdgp import MFDGP
    x, ys, fid = forrester_data()
    model = MFDGP(x, ys["obj1"], fid, 2)
    keys = set(model.state_dict().keys())
    for k in ("hidden_layer_0.covar_module.raw_outputscale",
              "hidden_layer_0.covar_module.base_kernel.raw_lengthscale",
              "hidden_layer_1.covar_module.kernels.0.kernels.0.raw_outputscale",
              "hidden_layer_1.covar_module.kernels.0.kernels.1.kernels.0.raw_variance",
              "hidden_layer_1.covar_module.kernels.0.kernels.1.kernels.1.base_kernel.raw_lengthscale",
              "hidden_layer_1.covar_module.kernels.1.base_kernel.raw_lengthscale",
              "hidden_layer_1.variational_strategy._variational_distribution.chol_variational_covar",
              "hidden_layer_1.variational_strategy.inducing_points",
              "hidden_layer_likelihood_1.noise_covar.raw_noise",
              # quirk Q12: the previous layer is a registered sub-module of the strategy
              "hidden_layer_1.variational_strategy.previous_layer.covar_module.raw_outputscale"):
        assert k in keys, k
    # modules() de-duplicates: each parameter once
    assert len(list(model.parameters())) == len(set(id(p) for p in model.parameters()))
    import copy
    m2 = copy.deepcopy(model)
    m2.load_state_dict(model.state_dict())


def test_median_lengthscale_order_statistic_equals_literal_formula():
    """The scalable form of get_init_lengthscale(TL.MEDIAN) is bit-identical to the reference's literal row-gather
    (quirk Q2) wherever the latter is computable."""
    import torch
    from mobocmf_b200.models.mfdgp import MFDGP
    g = torch.Generator().manual_seed(0)
    for n, d in [(2, 1), (3, 2), (12, 1), (16, 2), (37, 3), (60, 6)]:
        x = torch.rand(n, d, generator=g, dtype=torch.float64)
        fast = MFDGP.median_lengthscale(x)
        literal = MFDGP.median_lengthscale(x, literal=True)
        assert torch.equal(fast, literal), (n, d, float(fast), float(literal))
    # the Forrester fixture's derived known answers
    lf = torch.linspace(0, 1.0, 12, dtype=torch.float64)[:, None]
    hf = torch.tensor([0.1, 0.3, 0.5, 0.7], dtype=torch.float64)[:, None]
    assert abs(float(MFDGP.median_lengthscale(lf)) - 3.0 / 11.0) < 1e-12
    assert abs(float(MFDGP.median_lengthscale(hf)) - 0.2) < 1e-12
