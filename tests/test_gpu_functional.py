"""GPU parity of the C-ABI kernels (through mobocmf_b200.functional) against the CPU fp64 oracle."""
import math

import pytest
import torch

from oracle import mfdgp_oracle as O
from tests.helpers import adjudicate_state, random_state, clone_state, param_keys, relerr

pytestmark = pytest.mark.gpu

TOL = 1e-10   # fp64 parity bar of BASELINE.json's north_star, on well-conditioned inputs


def thetas(sd, L):
    """Constrained hyper-parameter vectors in the kernels' layout (include/mobocmf_b200.h)."""
    sp = torch.nn.functional.softplus
    out = []
    for l in range(L):
        c = "hidden_layer_%d.covar_module." % l
        if l == 0:
            out.append(torch.cat([sp(sd[c + "raw_outputscale"]).reshape(1),
                                  sp(sd[c + "base_kernel.raw_lengthscale"]).reshape(-1)]))
        else:
            out.append(torch.cat([
                sp(sd[c + "kernels.0.kernels.0.raw_outputscale"]).reshape(1),
                sp(sd[c + "kernels.0.kernels.1.kernels.0.raw_variance"]).reshape(1),
                sp(sd[c + "kernels.0.kernels.1.kernels.1.raw_outputscale"]).reshape(1),
                sp(sd[c + "kernels.0.kernels.1.kernels.1.base_kernel.raw_lengthscale"]).reshape(1),
                sp(sd[c + "kernels.1.raw_outputscale"]).reshape(1),
                sp(sd[c + "kernels.0.kernels.0.base_kernel.raw_lengthscale"]).reshape(-1),
                sp(sd[c + "kernels.1.base_kernel.raw_lengthscale"]).reshape(-1)]))
    return out


def cuda_forward(sd, L, x, eps, training=True):
    from mobocmf_b200 import functional as F
    th = thetas(sd, L)
    Zx = sd["hidden_layer_0.variational_strategy.inducing_points"]
    M = Zx.shape[0]
    outs, kls = [], []
    mu = var = None
    for l in range(L):
        q = "hidden_layer_%d.variational_strategy._variational_distribution." % l
        m, Lq = sd[q + "variational_mean"], sd[q + "chol_variational_covar"]
        zf = None if l == 0 else sd["hidden_layer_%d.variational_strategy._variational_distribution."
                                    "variational_mean" % (l - 1)]
        ops = F.layer_operators(th[l], zf, m, Lq, Zx, 0 if l == 0 else 1)
        kls.append(F.ops_kl(ops, M))
        if l == 0:
            mu, var = F.layer_rows(ops, th[l], None, Zx, x, kind=0, training=training)
        else:
            mu, var = F.layer_rows(ops, th[l], zf, Zx, x, mu_prev=mu, var_prev=var, eps=eps[l].reshape(-1), kind=1,
                                   training=training)
        outs.append((mu, var))
    return outs, kls


def torch_elbo(sd, L, noise_upper, outs, kls, y, fid, num_data):
    data = 0.0
    for l in range(L):
        noise = O.noise_value(sd["hidden_layer_likelihood_%d.noise_covar.raw_noise" % l], O.NOISE_LOWER,
                              noise_upper[l]).reshape(())
        mu, var = outs[l]
        ell = -0.5 * (((y - mu) ** 2 + var.clamp_min(1e-10)) / noise + noise.log() + math.log(2 * math.pi))
        data = data + (ell * (fid == l)).sum()
    B = y.shape[0]
    kl = sum(kls)
    return data - kl * B / num_data


# (M, d, R, L, lengthscale): small lengthscales give well-conditioned K_zz (the 1e-10 claim); the larger ones are
# the reference-default-like regime where both sides are only accurate to ~cond * eps (SURVEY.md §7 "Hard parts").
CASES = [(40, 2, 100, 2, 0.08), (256, 6, 300, 3, 0.3), (16, 1, 16, 2, 0.05), (75, 2, 75, 2, 0.06),
         (40, 2, 100, 2, 0.5), (256, 6, 130, 3, 1.0), (64, 3, 200, 2, 0.2)]


@pytest.mark.parametrize("M,d,R,L,ls", CASES)
def test_forward_backward_matches_oracle(M, d, R, L, ls):
    dev = torch.device("cuda:0")
    sd, noise_upper = random_state(M, d, L, seed=M + d, ls=ls)
    g = torch.Generator().manual_seed(7)
    x = torch.rand(R, d, generator=g, dtype=torch.float64)
    y = torch.randn(R, 1, generator=g, dtype=torch.float64)
    fid = torch.randint(0, L, (R, 1), generator=g).double()
    eps = [None] + [torch.randn(1, R, generator=g).double() for _ in range(1, L)]
    num_data = 3 * R

    # oracle (CPU, autograd)
    sdo = clone_state(sd, requires_grad=True)
    outs_o = O.mfdgp_forward(sdo, L, x, eps=eps, training=True)
    elbo_o, _ = O.elbo(sdo, L, noise_upper, outs_o, y.T, fid, num_data)
    elbo_o.backward()
    conds = [float(torch.linalg.cond(O.layer_kernel(sd, l, O.layer_inducing_points(sd, l), O.layer_inducing_points(sd, l))
                                     + 1e-6 * torch.eye(M, dtype=torch.float64))) for l in range(L)]

    # CUDA
    sdc = clone_state(sd, device=dev, requires_grad=True)
    outs_c, kls = cuda_forward(sdc, L, x.to(dev), [None if e is None else e.to(dev) for e in eps])
    elbo_c = torch_elbo(sdc, L, noise_upper, outs_c, kls, y.to(dev).reshape(-1), fid.to(dev).reshape(-1), num_data)
    elbo_c.backward()
    torch.cuda.synchronize()

    print("cond(P) per layer:", ["%.2e" % c for c in conds])
    tol = max(TOL, 20 * 2.2e-16 * max(conds))
    print("tolerance %.1e" % tol)
    for l in range(L):
        mo, vo = outs_o[l]
        mc, vc = outs_c[l]
        em, ev = relerr(mc, mo.reshape(-1)), relerr(vc, vo.reshape(-1))
        print("layer %d  mean relerr %.2e  var relerr %.2e" % (l, em, ev))
        assert em < tol and ev < tol
    kl_o = O.kl_divergence(sd, L)
    assert abs(float(sum(kls)) - float(kl_o)) / abs(float(kl_o)) < tol
    assert abs(float(elbo_c) - float(elbo_o)) / abs(float(elbo_o)) < tol
    worst = 0.0
    for k in param_keys(sd):
        go, gc = sdo[k].grad, sdc[k].grad
        assert gc is not None, k
        if "chol_variational_covar" in k:
            gc = torch.tril(gc)
            go = torch.tril(go)
        e = relerr(gc, go)
        worst = max(worst, e)
        print("grad %-90s relerr %.2e" % (k, e))
    if max(conds) < 1e5:
        assert worst < 100 * tol
    else:
        # ill-conditioned: the longdouble truth decides whether the CUDA gradients are as good as the oracle's
        keys = param_keys(sd)
        adjudicate_state(sd, O.NOISE_LOWER, noise_upper, -elbo_c.detach(), {k: -sdc[k].grad for k in keys},
                         -elbo_o.detach(), {k: -sdo[k].grad for k in keys}, L, x, y, fid, eps, num_data, 1)


@pytest.mark.parametrize("M,d,n,S,training", [(256, 6, 37, 64, True), (200, 3, 101, 25, False), (75, 2, 9, 11, True),
                                                (16, 1, 5, 40, True)])
def test_sample_tiled_rows_kernel_equals_generic_kernel(M, d, n, S, training):
    """Rows that are S MC samples of n points (xrep = S >= 11) run in the warp-specialised forward kernel
    (row_fwd_ws_kernel: product / build / finish warps); the same rows with x tiled explicitly (xrep = 1) run in the
    generic two-CTA kernel.  Same arithmetic per element, different summation order of the row sums: mean, variance and
    every gradient must agree to rounding, at ragged sizes (last tile partial, MP < 256, tiles spanning 2 - 3 points),
    and a second launch must reproduce the first bit for bit (no race in the barrier hand-overs)."""
    from mobocmf_b200 import functional as F
    L = 2
    sd, _ = random_state(M, d, L, seed=7, ls=0.3)
    sd = clone_state(sd, device="cuda:0")
    th = thetas(sd, L)
    Zx = sd["hidden_layer_0.variational_strategy.inducing_points"].contiguous()
    q = "hidden_layer_%d.variational_strategy._variational_distribution."
    zf = sd[(q % 0) + "variational_mean"]
    g = torch.Generator().manual_seed(3)
    x = torch.rand(n, d, generator=g, dtype=torch.float64).cuda()
    mu0 = torch.randn(n, generator=g, dtype=torch.float64).cuda()
    var0 = (0.05 + torch.rand(n, generator=g, dtype=torch.float64)).cuda()
    eps = torch.randn(n * S, generator=g, dtype=torch.float64).cuda()
    wmu, wvar = torch.randn(n * S, generator=g, dtype=torch.float64).cuda(), torch.randn(n * S, generator=g, dtype=torch.float64).cuda()

    def run(tiled):
        leaves = [t.detach().clone().requires_grad_(True) for t in (th[1], zf, sd[(q % 1) + "variational_mean"],
                                                                     sd[(q % 1) + "chol_variational_covar"], mu0, var0)]
        theta, zf_, m, Lq, mu_p, var_p = leaves
        ops = F.layer_operators(theta, zf_, m, Lq, Zx, 1)
        if tiled:     # generic kernel: explicit rows
            mu, var = F.layer_rows(ops, theta, zf_, Zx, x.repeat_interleave(S, 0).contiguous(),
                                   mu_prev=mu_p.repeat_interleave(S), var_prev=var_p.repeat_interleave(S), eps=eps, kind=1,
                                   training=training)
        else:         # warp-specialised kernel: row r reads x[r // S], mu_prev[r // S]
            mu, var = F.layer_rows(ops, theta, zf_, Zx, x, mu_prev=mu_p, var_prev=var_p, eps=eps, kind=1, xrep=S, prep=S,
                                   eps_mod=n * S, R=n * S, training=training)
        loss = (wmu * mu).sum() + (wvar * var).sum()
        grads = torch.autograd.grad(loss, leaves)
        return mu.detach(), var.detach(), [gr.detach() for gr in grads]

    a, b, c = run(False), run(True), run(False)
    assert torch.equal(a[0], c[0]) and torch.equal(a[1], c[1])
    for ga, gc in zip(a[2], c[2]):
        assert torch.equal(ga, gc)
    assert relerr(a[0], b[0]) < 1e-12 and relerr(a[1], b[1]) < 1e-12, (relerr(a[0], b[0]), relerr(a[1], b[1]))
    for k, (ga, gb) in enumerate(zip(a[2], b[2])):
        assert relerr(ga, gb) < 1e-10, (k, relerr(ga, gb))
