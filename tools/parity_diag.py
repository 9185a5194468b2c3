"""Development tool: where does the CUDA path lose digits on an ill-conditioned layer?  Runs the layer-0 operator chain
and row pass of the Forrester model (cond(K_zz + jitter I) ~ 1e7) through the C-ABI kernels and compares every
intermediate (P, L, W, H, beta, mean, variance, the gradients of an ELL-shaped loss) with numpy longdouble, in units of
eps * cond.   python tools/parity_diag.py   (needs a GPU)"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
LD = np.longdouble
EPS = 2.220446049250313e-16


def chol(A):
    n = A.shape[0]
    L = np.zeros_like(A)
    for j in range(n):
        L[j, j] = np.sqrt(A[j, j] - np.dot(L[j, :j], L[j, :j]))
        for i in range(j + 1, n):
            L[i, j] = (A[i, j] - np.dot(L[i, :j], L[j, :j])) / L[j, j]
    return L


def inv_lower(L):
    n = L.shape[0]
    W = np.zeros_like(L)
    for c in range(n):
        for i in range(c, n):
            W[i, c] = ((1 if i == c else 0) - np.dot(L[i, c:i], W[c:i, c])) / L[i, i]
    return W


def reference(dt, Z, X, a, l, m, Lq, y, noise, g, jit=1e-6):
    """The whitened formulation of csrc/matrix_ops.cu in dtype dt: values, gradients wrt (a, l), intermediates."""
    Z, X, m, y = Z.astype(dt), X.astype(dt), m.astype(dt), y.astype(dt)
    a, l, g, jit, noise = dt(a), dt(l), dt(g), dt(jit), dt(noise)
    Lq = np.tril(Lq.astype(dt))
    D2 = lambda A, B: ((A[:, None] - B[None, :]) / l) ** 2
    M = len(Z)
    Ezz, Ezx = np.exp(-D2(Z, Z) / 2), np.exp(-D2(Z, X) / 2)
    P = a * Ezz + jit * np.eye(M, dtype=dt)
    L = chol(P)
    W = inv_lower(L)
    beta, H = W @ m, W @ Lq
    K = a * Ezx
    T = W @ K
    U = H.T @ T
    mu, q1, q2 = beta @ T, (T * T).sum(0), (U * U).sum(0)
    c = a - q1
    mask = (c >= 0).astype(dt)
    var = np.maximum(c, 0) + q2
    kl = 0.5 * (2 * np.log(np.diag(L)).sum() - np.log(np.diag(Lq) ** 2).sum() + beta @ beta + (H * H).sum() - M)
    loss = (0.5 * ((y - mu) ** 2 + var) / noise).sum() + g * kl
    dmu, dvar = -(y - mu) / noise, 0.5 / noise * np.ones_like(mu)
    A2, A1, b, G2 = (T * dvar) @ T.T, (T * (dvar * mask)) @ T.T, T @ dmu, H @ H.T
    V = A2 @ G2
    I = np.eye(M, dtype=dt)
    N = A1 - V - V.T - 0.5 * (np.outer(beta, b) + np.outer(b, beta)) + 0.5 * g * (I - np.outer(beta, beta) - G2)
    dP = W.T @ N @ W
    dT = dmu[None, :] * beta[:, None] - 2 * dvar[None, :] * (mask[None, :] * T - H @ U)
    dK = W.T @ dT
    ga_zz, ga_rows, ga_xx = (dP * Ezz).sum(), (dK * Ezx).sum(), (dvar * mask).sum()
    gl_zz, gl_rows = (dP * a * Ezz * D2(Z, Z)).sum() / l, (dK * K * D2(Z, X)).sum() / l
    return dict(P=P, L=L, W=W, H=H, beta=beta, mu=mu, var=var, kl=kl, loss=loss, ga_zz=ga_zz, ga_rows=ga_rows + ga_xx,
                gl_zz=gl_zz, gl_rows=gl_rows, ga=ga_zz + ga_rows + ga_xx, gl=gl_zz + gl_rows)


def main():
    from mobocmf_b200 import functional as F
    from tests.helpers import forrester_data
    from tests.test_gpu_model import build
    dev = "cuda:0"
    x, ys, fid = forrester_data()
    model = build(x, ys["obj1"], fid, 2)
    lay = model.hidden_layer_0
    g = torch.Generator().manual_seed(5)
    perm = torch.randperm(x.shape[0], generator=g)
    xb, yb = x[perm].contiguous(), ys["obj1"][perm].reshape(-1).contiguous()
    vd = lay.variational_strategy._variational_distribution
    theta = lay.theta().detach().clone().requires_grad_(True)
    m_, Lq_ = vd.variational_mean.detach().clone(), vd.chol_variational_covar.detach().clone()
    Zx = lay._Zx().detach()
    M = Zx.shape[0]
    noise, gkl = 1e-3, 1.0
    ops = F.layer_operators(theta, None, m_, Lq_, Zx, 0, 1e-6)
    theta_rows = theta.detach().clone().requires_grad_(True)       # separate leaf: the row-pass share of the gradient
    mu, var = F.layer_rows(ops, theta_rows, None, Zx, xb.to(dev), kind=0, training=True)
    yd = yb.to(dev)
    loss = (0.5 * ((yd - mu) ** 2 + var) / noise).sum() + gkl * F.ops_kl(ops, M)
    loss.backward()
    lo = F.ops_layout(M)
    MP = lo["MP"]
    o = ops.detach().cpu().numpy()
    blk = lambda k: o[lo[k]:lo[k] + MP * MP].reshape(MP, MP)[:M, :M]
    th = theta.detach().cpu().numpy()
    args = (Zx.cpu().numpy()[:, 0], xb.numpy()[:, 0], th[0], th[1], m_.cpu().numpy(), Lq_.cpu().numpy(), yb.numpy(), noise, gkl)
    t, r = reference(LD, *args), reference(np.float64, *args)
    cond = np.linalg.cond(t["P"].astype(np.float64))
    print("cond %.2e   (errors below in units of eps * cond, relative to max |truth|; numpy fp64 of the same formulation | CUDA)" % cond)

    def rel(v, tv):
        tv = np.asarray(tv, dtype=LD)
        return float(np.max(np.abs(np.asarray(v, dtype=LD) - tv)) / np.max(np.abs(tv))) / (EPS * cond)
    cuda = dict(P=blk("P"), L=blk("L"), W=blk("W"), H=blk("H"), beta=o[lo["beta"]:lo["beta"] + M],
                mu=mu.detach().cpu().numpy(), var=var.detach().cpu().numpy(), kl=o[lo["scal"]], loss=float(loss),
                ga_zz=float(theta.grad[0]), ga_rows=float(theta_rows.grad[0]), gl_zz=float(theta.grad[1]),
                gl_rows=float(theta_rows.grad[1]))
    cuda["ga"], cuda["gl"] = cuda["ga_zz"] + cuda["ga_rows"], cuda["gl_zz"] + cuda["gl_rows"]
    for k in ("P", "L", "W", "H", "beta", "mu", "var", "kl", "loss", "ga_zz", "ga_rows", "ga", "gl_zz", "gl_rows", "gl"):
        print("%-8s truth max %.6e   numpy fp64 %8.3f   CUDA %8.3f" % (k, float(np.max(np.abs(t[k]))), rel(r[k], t[k]),
                                                                        rel(cuda[k], t[k])))


def full_model():
    """Second section: the 2-layer Forrester step.  Is the error of layer 0's hyper-parameter gradients made in layer 0's
    own backward, or inherited from d loss / d (mu_0, var_0), i.e. from layer 1's gradient wrt its propagated input?"""
    from mobocmf_b200.gp import settings
    from mobocmf_b200.mlls.variational_elbo_mf import VariationalELBOMF
    from oracle import mfdgp_oracle as O
    from oracle import mfdgp_truth as T
    from tests.helpers import forrester_data, oracle_view
    from tests.test_gpu_model import build
    dev = "cuda:0"
    x, ys, fid = forrester_data()
    y, L = ys["obj1"], 2
    model = build(x, y, fid, L)
    N = x.shape[0]
    elbo = VariationalELBOMF(model, N, L)
    g = torch.Generator().manual_seed(5)
    perm = torch.randperm(N, generator=g)
    xb, yb, fb = x[perm], y[perm], fid[perm]
    eps = [None] + [torch.randn(1, N, generator=g) for _ in range(1, L)]
    with settings.num_likelihood_samples(1):
        out = model(xb.to(dev), eps=[None if e is None else e.to(dev) for e in eps])
        out[0].mean.retain_grad(); out[0].raw_variance.retain_grad()
        res = elbo(out, yb.to(dev).T, fb.to(dev))
    (-res[0]).backward()
    sd, lo, up, _ = oracle_view(model)
    names = [n for n, _ in model.named_parameters()]
    for n in names:
        sd[n].requires_grad_(True)
    outs = O.mfdgp_forward(sd, L, xb, eps=eps, training=True)
    outs[0][0].retain_grad(); outs[0][1].retain_grad()
    e, kl = O.elbo(sd, L, up, outs, yb.T, fb, N, noise_lower=lo)
    (-e).backward(retain_graph=True)
    r = lambda a, b: float((a.detach().cpu() - b.detach().cpu()).abs().max() / b.detach().abs().max())
    print("\nfull model: CUDA vs fp64 oracle")
    print("  mu_0 %.2e  var_0 %.2e  dL/dmu_0 %.2e  dL/dvar_0 %.2e" % (
        r(out[0].mean, outs[0][0]), r(out[0].raw_variance, outs[0][1]), r(out[0].mean.grad, outs[0][0].grad),
        r(out[0].raw_variance.grad, outs[0][1].grad)))
    p0 = [n for n in names if n.startswith("hidden_layer_0.")]
    _, _, g_t = T.elbo_step_truth({k: v.detach() for k, v in sd.items()}, names, L, up, xb, yb, fb, eps, N, 1, noise_lower=lo)
    # oracle Jacobian of layer 0 applied to CUDA's incoming gradients (plus the oracle's own direct KL share)
    own = torch.autograd.grad([outs[0][0], outs[0][1]], [sd[n] for n in p0],
                              [outs[0][0].grad, outs[0][1].grad], retain_graph=True, allow_unused=True)
    mixed = torch.autograd.grad([outs[0][0], outs[0][1]], [sd[n] for n in p0],
                                [out[0].mean.grad.cpu(), out[0].raw_variance.grad.cpu()], retain_graph=True,
                                allow_unused=True)
    for n, a_, b_ in zip(p0, own, mixed):
        if a_ is None:
            continue
        gc = dict(model.named_parameters())[n].grad.cpu()
        go = sd[n].grad
        gm = go - a_ + b_                      # oracle gradient with CUDA's d loss / d (mu_0, var_0) substituted
        if "chol" in n:
            gc, go, gm = torch.tril(gc), torch.tril(go), torch.tril(gm)
        gt = np.tril(g_t[n]) if "chol" in n else g_t[n]
        print("  %-85s |cuda-truth| %.2e  |oracle-truth| %.2e  |oracle(J) x cuda(incoming)-truth| %.2e" % (
            n, T.err_vs(gc, gt), T.err_vs(go, gt), T.err_vs(gm, gt)))


if __name__ == "__main__":
    main()
    full_model()
