"""GPU edge cases of the row / operator kernels against the oracle: ragged row counts (1 row, tile size +- 1), M not a
multiple of 32 and below one block, d from 1 to 8 (the ARD limit), a single-fidelity model, empty candidate batches,
shape limits returning errors instead of wrong numbers."""
import pytest
import torch

from oracle import mfdgp_oracle as O
from tests.helpers import random_state, clone_state, param_keys, relerr, model_from_state

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _cond(sd, L):
    worst = 1.0
    for l in range(L):
        Z = O.layer_inducing_points(sd, l)
        P = O.layer_kernel(sd, l, Z, Z) + 1e-6 * torch.eye(Z.shape[0], dtype=torch.float64)
        worst = max(worst, float(torch.linalg.cond(P)))
    return worst


@pytest.mark.parametrize("M,d,L,B,S", [(5, 1, 2, 1, 1), (33, 8, 2, 31, 1), (32, 3, 1, 33, 1), (17, 2, 3, 65, 2),
                                       (96, 4, 2, 16, 4), (2, 1, 2, 3, 1)])
def test_ragged_shapes_fused_step(M, d, L, B, S):
    from mobocmf_b200.fused import FusedELBOStep
    from mobocmf_b200.mlls.variational_elbo_mf import VariationalELBOMF
    sd, noise_upper = random_state(M, d, L, seed=M + d + B, ls=0.35 if d < 6 else 0.8)
    g = torch.Generator().manual_seed(B)
    x = torch.rand(B, d, generator=g, dtype=torch.float64)
    y = torch.randn(B, 1, generator=g, dtype=torch.float64)
    fid = torch.randint(0, L, (B, 1), generator=g).double()
    eps = [None] + [torch.randn(B * S, generator=g).double() for _ in range(1, L)]
    sdo = clone_state(sd, requires_grad=True)
    loss_o, kl_o = O.elbo_step_loss_tiled(sdo, L, noise_upper, x, y, fid, eps, 5 * B, S)
    loss_o.backward()
    model = model_from_state(sd, noise_upper, L)
    step = FusedELBOStep(model, VariationalELBOMF(model, 5 * B, L))
    loss, kl = step(x.to(DEV), y.to(DEV), fid.to(DEV), eps=[None if e is None else e.to(DEV) for e in eps],
                    num_samples=S)
    step.check()
    cond = _cond(sd, L)
    tol = max(1e-10, 20 * 2.2e-16 * cond)
    assert relerr(loss, loss_o) < tol and relerr(kl, kl_o) < tol, (relerr(loss, loss_o), tol)
    grads = {n: p.grad for n, p in model.named_parameters()}
    for k in param_keys(sd):
        gp, go = grads[k], sdo[k].grad
        if "chol_variational_covar" in k:
            gp, go = torch.tril(gp), torch.tril(go)
        if go is None or float(go.abs().max()) == 0.0:     # parameter not reached by this minibatch (e.g. one fidelity only)
            assert float(gp.abs().max()) < 1e-12, k
            continue
        assert relerr(gp.reshape(-1), go.reshape(-1)) < (1e3 * tol if cond < 1e5 else 1e-2), (k, relerr(gp.reshape(-1), go.reshape(-1)))


@pytest.mark.parametrize("n", [1, 31, 32, 33, 257])
def test_acquisition_ragged_candidate_counts(n):
    M, d, L = 20, 2, 3
    sd, noise_upper = random_state(M, d, L, seed=3, ls=0.4)
    g = torch.Generator().manual_seed(n)
    samples = [torch.randn(25, 1, generator=g) for _ in range(L)]
    model = model_from_state(sd, noise_upper, L, samples=samples)
    model.eval()
    X = torch.rand(n, 1, d, generator=g, dtype=torch.float64)
    tol = max(1e-10, 20 * 2.2e-16 * _cond(sd, L))
    for f in range(L):
        with torch.no_grad():
            mu, var = model.predict_for_acquisition(X.to(DEV), f)
        model.eval()
        mu_o, var_o = O.predict_for_acquisition(sd, L, noise_upper, samples, X, f)
        assert mu.shape == (n,) and relerr(mu, mu_o) < tol and relerr(var, var_o) < 10 * tol


def test_shape_limits_are_errors_not_wrong_numbers():
    from mobocmf_b200 import _lib
    lib = _lib.load()
    t = torch.zeros(16, dtype=torch.float64, device=DEV)
    p = _lib.ptr(t)
    # M > 256 and d > 8 are outside the kernels' limits
    rc = lib.mobo_layer_precompute(0, 2, 300, p, None, p, p, p, 1e-6, p, _lib.stream_ptr())
    assert rc == -2
    rc = lib.mobo_layer_rows_fwd(0, 9, 16, p, None, p, p, p, 1, None, None, 1, None, 1, None, 4, 1, p, p, None, None,
                                 None, None, _lib.stream_ptr())
    assert rc == -2
    # zero rows: nothing is launched, success
    rc = lib.mobo_layer_rows_fwd(0, 2, 16, p, None, p, p, p, 1, None, None, 1, None, 1, None, 0, 1, p, p, None, None,
                                 None, None, _lib.stream_ptr())
    assert rc == 0
    torch.cuda.synchronize()


def test_cpu_tensors_are_refused():
    from mobocmf_b200 import functional as F
    sd, _ = random_state(8, 2, 1, seed=1)
    Zx = sd["hidden_layer_0.variational_strategy.inducing_points"]
    with pytest.raises(RuntimeError):
        F.layer_operators(torch.ones(3, dtype=torch.float64), None, torch.zeros(8, dtype=torch.float64),
                          torch.eye(8, dtype=torch.float64), Zx, 0)


def test_not_positive_definite_is_reported():
    """Upstream raises NotPSDError from psd_safe_cholesky; the fused step reports it through its status word."""
    from mobocmf_b200.fused import FusedELBOStep
    from mobocmf_b200.mlls.variational_elbo_mf import VariationalELBOMF
    M, d, L, B = 24, 2, 2, 8
    sd, noise_upper = random_state(M, d, L, seed=2, ls=0.4)
    Zk = "hidden_layer_0.variational_strategy.inducing_points"
    sd[Zk][1] = sd[Zk][0]                      # duplicated inducing input ...
    sd["hidden_layer_1.variational_strategy.inducing_points"][:, :d] = sd[Zk]
    model = model_from_state(sd, noise_upper, L)
    for l in range(L):
        getattr(model, "hidden_layer_%d" % l).variational_strategy.jitter_val = -1e-3   # ... and a negative jitter
    step = FusedELBOStep(model, VariationalELBOMF(model, 40, L))
    g = torch.Generator().manual_seed(0)
    x = torch.rand(B, d, generator=g, dtype=torch.float64).to(DEV)
    step(x, torch.zeros(B, 1, dtype=torch.float64, device=DEV), torch.zeros(B, 1, dtype=torch.float64, device=DEV))
    with pytest.raises(RuntimeError, match="NotPSDError|NanError"):
        step.check()
