"""GPU parity of the fused ELBO step (mobo_elbo_step), the Adam kernel (mobo_adam) and the fused acquisition chain
(mobo_acq_moments / mobo_jes) against the CPU oracle and against the composable autograd path."""
import copy

import pytest
import torch

from oracle import mfdgp_oracle as O
from tests.helpers import adjudicate_step, synthetic_data, forrester_data, oracle_view, relerr, parity_tol

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def make_model(L=3, M=48, ls=0.25, seed=1, n_per=(60, 40, 20), d=3):
    from mobocmf_b200.models.mfdgp import MFDGP
    x, y, fid = synthetic_data(list(n_per[:L]), d, seed=3)
    N = x.shape[0]
    torch.manual_seed(seed)
    perm = torch.randperm(N)
    x, y, fid = x[perm], y[perm], fid[perm]
    model = MFDGP(x, y, fid, L, num_inducing=M, init_lengthscale=ls)
    model.double()
    g = torch.Generator().manual_seed(seed + 5)
    with torch.no_grad():     # move off the initial point so that every term matters
        for n, p in model.named_parameters():
            if "chol_variational_covar" in n:
                p.add_(torch.tril(torch.randn(p.shape, generator=g, dtype=p.dtype)) * 0.02 / p.shape[0] ** 0.5)
                p.diagonal().abs_().add_(0.05)
            else:
                p.add_(0.1 * torch.randn(p.shape, generator=g, dtype=p.dtype))
    return model.to(DEV), x, y, fid, N


@pytest.mark.parametrize("L,S,B,M,freeze", [(3, 5, 33, 48, None), (2, 1, 40, 32, None), (3, 1, 70, 75, "phase1"),
                                            (3, 4, 64, 64, "cond"), (2, 3, 130, 256, None),
                                            # 10 000 ragged rows in the upper layer: every persistent CTA of the
                                            # product / SYRK kernels walks several tiles (prefetch pipelines, mbarrier
                                            # phase wrap, fragment rings reused across tiles)
                                            (2, 40, 250, 256, None)])
def test_fused_step_matches_oracle_and_composable(L, S, B, M, freeze):
    from mobocmf_b200.fused import FusedELBOStep
    from mobocmf_b200.gp import settings
    from mobocmf_b200.mlls.variational_elbo_mf import VariationalELBOMF
    n_per = (300, 200, 100) if M > 100 else (60, 40, 20)
    model, x, y, fid, N = make_model(L=L, M=M, n_per=n_per)
    if freeze == "phase1":
        model.fix_variational_hypers(True)
    elif freeze == "cond":
        model.fix_variational_hypers_cond(True)
    elbo = VariationalELBOMF(model, N, L)
    g = torch.Generator().manual_seed(2)
    idx = torch.randint(0, N, (B,), generator=g)
    eps = [None] + [torch.randn(B * S, generator=g).double() for _ in range(1, L)]
    xb, yb, fb = x[idx].to(DEV), y[idx].to(DEV), fid[idx].to(DEV)
    eps_d = [None if e is None else e.to(DEV) for e in eps]

    tol, cond = parity_tol(model)
    step = FusedELBOStep(model, elbo)
    loss, kl = step(xb, yb, fb, eps=eps_d, num_samples=S)
    step.check()
    loss, kl = loss.clone(), kl.clone()
    g_fused = {n: (None if p.grad is None else p.grad.clone()) for n, p in model.named_parameters()}

    # composable autograd path, same kernels
    for p in model.parameters():
        p.grad = None
    with settings.num_likelihood_samples(1):
        out = model(xb, eps=eps_d, num_samples=S)
        res = elbo(out, yb.T, fb)
    (-res[0]).backward()
    assert relerr(loss, -res[0]) < 1e-12 and relerr(kl, res[1]) < 1e-12
    for n, p in model.named_parameters():
        if not p.requires_grad:
            assert g_fused[n] is None, n
            continue
        gf, gc = g_fused[n], p.grad
        if "chol_variational_covar" in n:
            gf, gc = torch.tril(gf), torch.tril(gc)
        # same kernels, different summation order of the partial gradients (the S samples of a point are folded by
        # ell_kernel here, by torch reductions there): equal up to cond * eps.  tol = 20 eps cond; measured up to 240 eps cond
        # on cancelled scalars (the v_lin gradient at cond 2e8); the bar against the ORACLE below is 1e3 tol
        assert relerr(gf, gc) < max(1e-9, 30 * tol), (n, relerr(gf, gc), cond)

    # oracle
    sd, lo, up, _ = oracle_view(model)
    names = [n for n, p in model.named_parameters() if p.requires_grad]
    for n in names:
        sd[n].requires_grad_(True)
    loss_o, kl_o = O.elbo_step_loss_tiled(sd, L, up, x[idx], y[idx], fid[idx], eps, N, S, noise_lower=lo)
    loss_o.backward()
    print("cond %.2e tol %.1e loss relerr %.2e" % (cond, tol, relerr(loss, loss_o)))
    assert relerr(loss, loss_o) < tol and relerr(kl, kl_o) < tol
    rows = B + (L - 1) * B * S
    if cond >= 1e5 and rows <= 2500:
        # ill-conditioned: fp64 itself loses digits, so the judge is the longdouble truth, not a wider bar
        adjudicate_step(model, loss, g_fused, loss_o.detach(), {n: sd[n].grad for n in names}, L, x[idx], y[idx],
                        fid[idx], eps, N, S)
        return
    if cond >= 1e5:
        # too many rows for the longdouble truth (minutes).  The per-row arithmetic of these kernels IS adjudicated by
        # the smaller cases above; what this case adds is several tiles per persistent CTA.  That is checked without
        # any conditioning caveat by additivity: the step's gradients equal the sum over 5 row shards, each small
        # enough for one tile per CTA (the single-tile path the truth has verified).
        full = torch.cat([g_fused[n].reshape(-1) for n, p in model.named_parameters() if p.requires_grad])
        acc = torch.zeros_like(full)
        nsh = 5
        for r in range(nsh):
            a, b = r * B // nsh, (r + 1) * B // nsh
            e = [None] + [t[a * S:b * S] for t in eps_d[1:]]
            step(xb[a:b], yb[a:b], fb[a:b], eps=e, num_samples=S)
            # (per parameter: the flat buffer leads with the d L_q blocks, not in named_parameters order)
            acc += torch.cat([p.grad.reshape(-1) for n, p in model.named_parameters() if p.requires_grad])
        # every shard carries the KL term scaled by its own B_r / N: they add up to the full step's B / N
        # (the SYRK partials fold in another order per partition: ~1e-16 sqrt(rows) there, times cond through
        # dP = W^T N W; an indexing slip in the multi-tile pipelines is an O(1) error)
        assert relerr(acc, full) < tol, (relerr(acc, full), tol)
    for n in names:
        gf, go = g_fused[n], sd[n].grad
        if "chol_variational_covar" in n:
            gf, go = torch.tril(gf), torch.tril(go)
        assert relerr(gf, go) < 1e3 * tol, (n, relerr(gf, go))


def test_fused_step_single_sample_matches_reference_shaped_oracle():
    """S = 1 with (1, B) float32 normals: exactly the reference's step (fitter.py:161-168) on the Forrester data."""
    from mobocmf_b200.fused import FusedELBOStep
    from mobocmf_b200.mlls.variational_elbo_mf import VariationalELBOMF
    from mobocmf_b200.models.mfdgp import MFDGP
    x, ys, fid = forrester_data()
    torch.manual_seed(0)
    model = MFDGP(x, ys["obj1"], fid, 2)
    model.double().to(DEV)
    N = 16
    elbo = VariationalELBOMF(model, N, 2)
    g = torch.Generator().manual_seed(5)
    perm = torch.randperm(N, generator=g)
    perm = torch.roll(perm, 1) if torch.equal(perm, torch.arange(N)) else perm
    xb, yb, fb = x[perm], ys["obj1"][perm], fid[perm]
    eps = [None, torch.randn(1, N, generator=g)]
    step = FusedELBOStep(model, elbo)
    loss, kl = step(xb.to(DEV), yb.to(DEV), fb.to(DEV), eps=[None, eps[1].to(DEV)])
    sd, lo, up, _ = oracle_view(model)
    loss_o, kl_o = O.elbo_step_loss(sd, 2, up, xb, yb, fb, eps, N, noise_lower=lo)
    tol, cond = parity_tol(model)
    assert relerr(loss, loss_o) < tol and relerr(kl, kl_o) < tol, (relerr(loss, loss_o), tol)


def test_adam_kernel_matches_torch_adam():
    from mobocmf_b200.fused import Adam
    g = torch.Generator().manual_seed(0)
    shapes = [(256, 256), (256,), (1, 6), (), (1,)]
    p_ref = [torch.randn(s, generator=g, dtype=torch.float64).to(DEV).requires_grad_(True) for s in shapes]
    p_our = [p.detach().clone().requires_grad_(True) for p in p_ref]
    o_ref = torch.optim.Adam([{"params": p_ref}], lr=0.003)
    o_our = Adam([{"params": p_our}], lr=0.003)
    for it in range(4):
        for a, b in zip(p_ref, p_our):
            gr = torch.randn(a.shape, generator=g, dtype=torch.float64).to(DEV) * (10.0 ** (it - 2))
            a.grad = gr.clone()
            b.grad = gr.clone()
        o_ref.step()
        o_our.step()
        for a, b in zip(p_ref, p_our):
            assert relerr(b, a) < 1e-14, (it, a.shape, relerr(b, a))
    sd = o_our.state_dict()
    assert set(sd["state"][0].keys()) == {"step", "exp_avg", "exp_avg_sq"}


def test_two_shards_sum_to_one_step():
    """Multi-GPU semantics on one device: the summed gradients of two half-batch steps equal the full-batch step."""
    from mobocmf_b200.fused import FusedELBOStep
    from mobocmf_b200.mlls.variational_elbo_mf import VariationalELBOMF
    from mobocmf_b200.util.distributed import shard_bounds
    L, S, B = 3, 2, 48
    model, x, y, fid, N = make_model(L=L, M=32)
    elbo = VariationalELBOMF(model, N, L)
    g = torch.Generator().manual_seed(4)
    idx = torch.randint(0, N, (B,), generator=g)
    eps = [None] + [torch.randn(B * S, generator=g).double().to(DEV) for _ in range(1, L)]
    xb, yb, fb = x[idx].to(DEV), y[idx].to(DEV), fid[idx].to(DEV)
    step = FusedELBOStep(model, elbo)
    loss, _ = step(xb, yb, fb, eps=eps, num_samples=S)
    full_loss = loss.clone()
    full = step.flat.flat.clone()
    acc = torch.zeros_like(full)
    tot = 0.0
    for r in range(2):
        lo, hi = shard_bounds(B, r, 2)
        e = [None] + [t[lo * S:hi * S] for t in eps[1:]]
        l_r, _ = step(xb[lo:hi], yb[lo:hi], fb[lo:hi], eps=e, num_samples=S)
        tot = tot + float(l_r)
        acc += step.flat.flat
    assert relerr(acc, full) < 1e-11, relerr(acc, full)
    assert abs(tot - float(full_loss)) < 1e-11 * abs(float(full_loss))


def test_training_loop_decreases_loss_and_matches_composable_loop():
    """A few fused steps + Adam kernel == the same steps through autograd + torch.optim.Adam."""
    from mobocmf_b200.fused import Adam, FusedELBOStep
    from mobocmf_b200.gp import settings
    from mobocmf_b200.mlls.variational_elbo_mf import VariationalELBOMF
    L, S, B = 2, 2, 32
    m1, x, y, fid, N = make_model(L=L, M=32)
    m2 = copy.deepcopy(m1)
    e1, e2 = VariationalELBOMF(m1, N, L), VariationalELBOMF(m2, N, L)
    o1 = Adam([{"params": m1.parameters()}], lr=0.01)
    o2 = torch.optim.Adam([{"params": m2.parameters()}], lr=0.01)
    step = FusedELBOStep(m1, e1)
    g = torch.Generator().manual_seed(8)
    losses = []
    idx = torch.randint(0, N, (B,), generator=g)       # a fixed minibatch and fixed normals: the loss must go down
    eps = [None] + [torch.randn(B * S, generator=g).double().to(DEV) for _ in range(1, L)]
    xb, yb, fb = x[idx].to(DEV), y[idx].to(DEV), fid[idx].to(DEV)
    for it in range(5):
        l1, _ = step(xb, yb, fb, eps=eps, num_samples=S)
        o1.step()
        o2.zero_grad()
        with settings.num_likelihood_samples(1):
            res = e2(m2(xb, eps=eps, num_samples=S), yb.T, fb)
        (-res[0]).backward()
        o2.step()
        losses.append(float(l1))
        assert relerr(l1, -res[0]) < 1e-9, (it, relerr(l1, -res[0]))
    for (n, a), (_, b) in zip(m1.named_parameters(), m2.named_parameters()):
        assert relerr(a, b) < 1e-8, (n, relerr(a, b))
    assert losses[-1] < losses[0]


# second case: 420 candidates x 25 samples = 10 500 rows per upper layer at M = 256: the persistent CTAs of the forward
# kernel walk several tiles in eval mode
@pytest.mark.parametrize("M,d,ncand,n_per", [(48, 2, 203, (60, 40, 20)), (256, 6, 420, (300, 200, 100))])
def test_fused_acquisition_chain_matches_python_path_and_oracle(M, d, ncand, n_per):
    from mobocmf_b200.acquisition_functions.JESMOC_MFDGP import _JES_MFDGP
    L = 3
    mu_model, x, y, fid, N = make_model(L=L, M=M, d=d, n_per=n_per)
    mc_model = copy.deepcopy(mu_model)
    g = torch.Generator().manual_seed(9)
    with torch.no_grad():
        for n, p in mc_model.named_parameters():
            if "chol_variational_covar" in n:
                p.mul_(0.6)
    X = torch.rand(ncand, 1, d, generator=g, dtype=torch.float64)
    for fidelity in range(L):
        acq = _JES_MFDGP(fidelity, mu_model, mc_model)
        with torch.no_grad():
            val = acq(X.to(DEV))                              # fused chain (no grad needed)
        Xg = X.to(DEV).requires_grad_(True)
        val_g = acq(Xg)                                       # composable chain (dX for optimize_acqf)
        assert relerr(val, val_g) < 1e-12
        mods = []
        for m in (mu_model, mc_model):
            sd, lo, up, samples = oracle_view(m)
            mods.append(dict(sd=sd, num_layers=L, noise_upper=up, noise_lower=lo, samples=samples))
        val_o = O.jes_mfdgp(mods[0], mods[1], X, fidelity)
        tol = max(parity_tol(mu_model)[0], parity_tol(mc_model)[0])
        assert float(val.max()) > 1e-3
        assert (val.cpu() - val_o).abs().max() < 100 * tol


def test_cuda_graph_replay_matches_eager_steps():
    """The captured (fused step + Adam) graph replays to the same parameters as the eager fused loop (Forrester-sized:
    the launch-latency-bound regime of the reference's own examples)."""
    from mobocmf_b200.fused import Adam, FusedELBOStep, GraphedELBOStep
    from mobocmf_b200.mlls.variational_elbo_mf import VariationalELBOMF
    L, S, B = 2, 1, 24
    m1, x, y, fid, N = make_model(L=L, M=24, n_per=(14, 10), d=2)
    m2 = copy.deepcopy(m1)
    e1, e2 = VariationalELBOMF(m1, N, L), VariationalELBOMF(m2, N, L)
    o1 = Adam([{"params": m1.parameters()}], lr=0.01, capturable=True)
    o2 = Adam([{"params": m2.parameters()}], lr=0.01)
    gstep = GraphedELBOStep(FusedELBOStep(m1, e1), o1, B, num_samples=S, static_eps=True)
    estep = FusedELBOStep(m2, e2)
    g = torch.Generator().manual_seed(3)
    for it in range(6):
        perm = torch.randperm(N, generator=g)[:B]
        eps = [None] + [torch.randn(B * S, generator=g).double().to(DEV) for _ in range(1, L)]
        xb, yb, fb = x[perm].to(DEV), y[perm].to(DEV), fid[perm].to(DEV)
        l1, _ = gstep(xb, yb, fb, eps=eps)
        l2, _ = estep(xb, yb, fb, eps=eps, num_samples=S)
        o2.step()
        assert relerr(l1, l2) < 1e-12, (it, relerr(l1, l2))
    for (n, a), (_, b) in zip(m1.named_parameters(), m2.named_parameters()):
        assert relerr(a, b) < 1e-12, (n, relerr(a, b))
    assert int(o1._step_dev) == 6 and len(o1._cohorts) == 1


def test_graphed_steps_with_ragged_last_batch_share_one_fused_step():
    """N % batch_size != 0 with CUDA graphs (the reference's DataLoader yields the short last batch like any other):
    one FusedELBOStep serves TWO captured graphs (B = 10 and B = 6).  Each graph bakes its workspace address in, so
    the workspaces must both stay alive; and capturing the second graph mid-training must not disturb the optimiser
    state accumulated with the first.  Three epochs of alternating replays equal the eager fused loop to 1e-12."""
    from mobocmf_b200.fused import Adam, FusedELBOStep, GraphedELBOStep
    from mobocmf_b200.mlls.variational_elbo_mf import VariationalELBOMF
    L, S = 2, 2
    m1, x, y, fid, N = make_model(L=L, M=16, n_per=(10, 6), d=2)
    m2 = copy.deepcopy(m1)
    e1, e2 = VariationalELBOMF(m1, N, L), VariationalELBOMF(m2, N, L)
    o1 = Adam([{"params": m1.parameters()}], lr=0.01, capturable=True)
    o2 = Adam([{"params": m2.parameters()}], lr=0.01)
    f1, f2 = FusedELBOStep(m1, e1), FusedELBOStep(m2, e2)
    graphs = {}
    g = torch.Generator().manual_seed(3)
    for epoch in range(3):
        perm = torch.randperm(N, generator=g)
        for a in (0, 10):
            idx = perm[a:a + 10]
            B = idx.numel()
            eps = [None] + [torch.randn(B * S, generator=g).double().to(DEV) for _ in range(1, L)]
            xb, yb, fb = x[idx].to(DEV), y[idx].to(DEV), fid[idx].to(DEV)
            if B not in graphs:
                graphs[B] = GraphedELBOStep(f1, o1, B, num_samples=S, static_eps=True)
            l1, _ = graphs[B](xb, yb, fb, eps=eps)
            l2, _ = f2(xb, yb, fb, eps=eps, num_samples=S)
            o2.step()
            assert relerr(l1, l2) < 1e-12, (epoch, a, relerr(l1, l2))
    assert set(graphs) == {10, 6} and len(f1._ws) == 2 and f1._ws_pinned == {(10, S), (6, S)}
    for (n, a), (_, b) in zip(m1.named_parameters(), m2.named_parameters()):
        assert relerr(a, b) < 1e-12, (n, relerr(a, b))
    assert int(o1._step_dev) == 6


def test_failed_step_is_skipped_by_adam_and_raised_at_check():
    """A step whose Cholesky cannot be rescued by the three jitter retries leaves the parameters untouched (the Adam
    kernel reads the step's failure flag on the device) and surfaces as NotPSDError at the next check()."""
    from mobocmf_b200.errors import NotPSDError
    from mobocmf_b200.fused import Adam, FusedELBOStep
    from mobocmf_b200.mlls.variational_elbo_mf import VariationalELBOMF
    L, S, B = 2, 1, 20
    model, x, y, fid, N = make_model(L=L, M=32)
    for l in range(L):
        getattr(model, "hidden_layer_%d" % l).variational_strategy.jitter_val = -1e-2
    step = FusedELBOStep(model, VariationalELBOMF(model, N, L))
    opt = Adam([{"params": model.parameters()}], lr=0.01)
    opt.skip_flag = step.skip_flag
    before = [p.detach().clone() for p in model.parameters()]
    g = torch.Generator().manual_seed(0)
    idx = torch.randint(0, N, (B,), generator=g)
    step(x[idx].to(DEV), y[idx].to(DEV), fid[idx].to(DEV))
    opt.step()
    for a, p in zip(before, model.parameters()):
        assert torch.equal(a, p.detach())
    assert step.retries() == 3
    with pytest.raises(NotPSDError):
        step.check()
    # with the regular jitter the same objects step normally again (the sticky flags were cleared by check())
    for l in range(L):
        getattr(model, "hidden_layer_%d" % l).variational_strategy.jitter_val = 1e-6
    step._sig = None
    step(x[idx].to(DEV), y[idx].to(DEV), fid[idx].to(DEV))
    opt.step()
    step.check()
    assert step.retries() == 0
    assert any(not torch.equal(a, p.detach()) for a, p in zip(before, model.parameters()))
