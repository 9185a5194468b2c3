"""GPU parity of the S > 1 training path (fact F4 extension) against the oracle's tiled restatement."""
import pytest
import torch

from oracle import mfdgp_oracle as O
from tests.helpers import synthetic_data, oracle_view, relerr, parity_tol

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_multisample_elbo_and_grads():
    from mobocmf_b200.gp import settings
    from mobocmf_b200.mlls.variational_elbo_mf import VariationalELBOMF
    from mobocmf_b200.models.mfdgp import MFDGP
    x, y, fid = synthetic_data([60, 40, 20], 3, seed=3)
    L, N, B, S, M = 3, 120, 33, 5, 48
    torch.manual_seed(1)
    perm = torch.randperm(N)
    x, y, fid = x[perm], y[perm], fid[perm]
    model = MFDGP(x, y, fid, L, num_inducing=M, init_lengthscale=0.25)
    model.double().to(DEV)
    elbo = VariationalELBOMF(model, N, L)
    g = torch.Generator().manual_seed(2)
    idx = torch.randint(0, N, (B,), generator=g)
    eps = [None] + [torch.randn(B * S, generator=g).double() for _ in range(1, L)]
    with settings.num_likelihood_samples(1):
        out = model(x[idx].to(DEV), eps=[None if e is None else e.to(DEV) for e in eps], num_samples=S)
        res = elbo(out, y[idx].to(DEV).T, fid[idx].to(DEV))
    (-res[0]).backward()
    sd, lo, up, _ = oracle_view(model)
    names = [n for n, _ in model.named_parameters()]
    for n in names:
        sd[n].requires_grad_(True)
    loss_o, kl_o = O.elbo_step_loss_tiled(sd, L, up, x[idx], y[idx], fid[idx], eps, N, S, noise_lower=lo)
    loss_o.backward()
    tol, cond = parity_tol(model)
    assert relerr(-res[0], loss_o) < tol and relerr(res[1], kl_o) < tol
    if cond >= 1e5:       # ill-conditioned: adjudicated by the longdouble truth (tests/helpers.adjudicate)
        from tests.helpers import adjudicate_step
        adjudicate_step(model, -res[0], {n: p.grad for n, p in model.named_parameters()}, loss_o.detach(),
                        {n: sd[n].grad for n in names}, L, x[idx], y[idx], fid[idx], eps, N, S)
        return
    for n, p in model.named_parameters():
        gp, go = p.grad, sd[n].grad
        if "chol_variational_covar" in n:
            gp, go = torch.tril(gp), torch.tril(go)
        assert relerr(gp, go) < 1e3 * tol, (n, relerr(gp, go))
