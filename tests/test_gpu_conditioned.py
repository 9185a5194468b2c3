"""GPU parity of the conditioned training step (util/blackbox_mfdgp_fitter.py:272-346) against the oracle."""
import pytest
import torch

from oracle import mfdgp_oracle as O
from tests.helpers import forrester_data, oracle_view, relerr, parity_tol

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_conditioned_step_loss_and_grads():
    from mobocmf_b200.util.blackbox_mfdgp_fitter import BlackBoxMFDGPFitter
    x, ys, fid = forrester_data()
    N, L = x.shape[0], 2
    torch.manual_seed(0)
    fitter = BlackBoxMFDGPFitter(L, N, num_epochs_1=0, num_epochs_2=0, device=torch.device(DEV))
    fitter.verbose = False
    fitter.initialize_mfdgp(x, ys["obj1"], fid, "obj1")
    fitter.initialize_mfdgp(x, ys["obj2"], fid, "obj2")
    fitter.initialize_mfdgp(x, ys["con1"], fid, "con1", threshold_constraint=0.1, is_constraint=True)
    g = torch.Generator().manual_seed(3)
    P, T = 7, 10
    fitter.pareto_set = torch.rand(P, 1, generator=g, dtype=torch.float64)
    fitter.pareto_front = torch.randn(P, 2, generator=g, dtype=torch.float64)
    x_tilde = torch.rand(T, 1, generator=g, dtype=torch.float64)
    hobjs = list(fitter.mfdgp_handlers_objs.values())
    hcons = list(fitter.mfdgp_handlers_cons.values())
    for h in hobjs + hcons:
        h.mfdgp.fix_variational_hypers_cond(True)
        with torch.no_grad():
            for n, p in h.mfdgp.named_parameters():
                if "variational_mean" in n:
                    p.add_(0.1 * torch.randn(p.shape, generator=g, dtype=p.dtype).to(DEV))
    batches, eps, eps_o = {}, {}, {}
    for key, h in [(("obj", i), h) for i, h in enumerate(hobjs)] + [(("con", k), h) for k, h in enumerate(hcons)]:
        perm = torch.randperm(N, generator=g)
        batches[key] = (h.x[perm.to(DEV)], h.y[perm.to(DEV)], h.f[perm.to(DEV)])
        e = {"batch": [None, torch.randn(1, N, generator=g)], "pareto": [None, torch.randn(1, P, generator=g)],
             "tilde": [None, torch.randn(1, T, generator=g)]}
        eps_o[key] = e
        eps[key] = {k: [None, v[1].to(DEV)] for k, v in e.items()}
    loss = fitter.conditioned_loss(hobjs, hcons, x_tilde=x_tilde.to(DEV), batches=batches, eps=eps)
    loss.backward()

    def omod(key, h):
        sd, lo, up, _ = oracle_view(h.mfdgp)
        for n, p in h.mfdgp.named_parameters():
            if p.requires_grad:
                sd[n].requires_grad_(True)
        xb, yb, fb = [t.cpu() for t in batches[key]]
        return dict(sd=sd, num_layers=L, noise_upper=up, num_data=N, batch=(xb, yb, fb),
                    eps_batch=eps_o[key]["batch"], eps_pareto=eps_o[key]["pareto"], eps_tilde=eps_o[key]["tilde"])
    objs = [omod(("obj", i), h) for i, h in enumerate(hobjs)]
    cons = [omod(("con", k), h) for k, h in enumerate(hcons)]
    loss_o = O.conditioned_step_loss(objs, cons, fitter.pareto_set, fitter.pareto_front, fitter.thresholds_cons,
                                     x_tilde, eps_factor=fitter.eps)
    loss_o.backward()
    tol = max(parity_tol(h.mfdgp)[0] for h in hobjs + hcons)
    assert relerr(loss, loss_o) < tol, (relerr(loss, loss_o), tol)
    for mod, h in zip(objs + cons, hobjs + hcons):
        for n, p in h.mfdgp.named_parameters():
            if not p.requires_grad:
                assert p.grad is None
                continue
            gp, go = p.grad, mod["sd"][n].grad
            if "chol_variational_covar" in n:
                gp, go = torch.tril(gp), torch.tril(go)
            assert relerr(gp, go) < 1e3 * tol, (n, relerr(gp, go), tol)


def test_graph_captured_conditioned_training_runs_and_learns():
    """The conditioned iteration captured in a CUDA graph (fitter(use_cuda_graph=True)): same loss function as the eager
    path (checked above against the oracle), so here: it replays, stays finite, lowers the loss like the eager loop
    does from the same start, and leaves the frozen parameters untouched."""
    from mobocmf_b200.util.blackbox_mfdgp_fitter import BlackBoxMFDGPFitter
    from tests.helpers import forrester_data
    x, ys, fid = forrester_data()
    finals = {}
    for use_graph in (False, True):
        torch.manual_seed(0)
        fitter = BlackBoxMFDGPFitter(2, 16, num_epochs_1=30, num_epochs_2=30, device=torch.device(DEV),
                                     use_cuda_graph=use_graph)
        fitter.verbose = False
        fitter.initialize_mfdgp(x, ys["obj1"], fid, "obj1")
        fitter.initialize_mfdgp(x, ys["obj2"], fid, "obj2")
        fitter.initialize_mfdgp(x, ys["con1"], fid, "con1", threshold_constraint=0.0, is_constraint=True)
        fitter.train_mfdgps()
        g = torch.Generator().manual_seed(1)
        fitter.pareto_set = torch.rand(7, 1, generator=g, dtype=torch.float64)
        fitter.pareto_front = torch.randn(7, 2, generator=g, dtype=torch.float64) * 0.3
        cond = fitter.copy_uncond()
        cond.pareto_set, cond.pareto_front = fitter.pareto_set, fitter.pareto_front
        cond.verbose = False
        noise_before = [float(h.mfdgp.hidden_layer_likelihood_1.noise_covar.raw_noise) for h in
                        cond.mfdgp_handlers_objs.values()]
        hs = (list(cond.mfdgp_handlers_objs.values()), list(cond.mfdgp_handlers_cons.values()))
        torch.manual_seed(5)
        with torch.no_grad():
            first = float(sum(cond.conditioned_loss(*hs) for _ in range(20)) / 20)
        cond.num_epochs_2 = 300
        cond.train_conditioned_mfdgps()
        torch.manual_seed(5)
        with torch.no_grad():
            last = float(sum(cond.conditioned_loss(*hs) for _ in range(20)) / 20)
        assert last < first, (use_graph, first, last)
        for hh in hs[0] + hs[1]:
            assert all(bool(torch.isfinite(p).all()) for p in hh.mfdgp.parameters())
        assert noise_before == [float(h.mfdgp.hidden_layer_likelihood_1.noise_covar.raw_noise) for h in hs[0]]
        if use_graph:
            assert cond._cond_graphs and next(iter(cond._cond_graphs.values())).graphs
        finals[use_graph] = (first, last)
    # the two loops start from the same point and optimise the same stochastic objective
    assert abs(finals[True][0] - finals[False][0]) < 1e-3 * abs(finals[False][0])   # same data, seeds; other RNG use
    assert abs(finals[True][1] - finals[False][1]) < 0.2 * abs(finals[False][0] - finals[False][1]) + 1.0


def test_graph_captured_conditioned_iteration_equals_eager():
    """The CUDA-graph replay of the conditioned iteration computes exactly what the eager iteration does (loss and, after
    several iterations, every parameter to 1e-12) when both get the same minibatches, x-tilde and training normals.
    The eager iteration is the one checked against the oracle above (which keeps torch.prod and ell[mask]); the captured
    one adds stream forks, static buffers and the device-resident Adam step count."""
    import copy
    from mobocmf_b200.fused import Adam
    from mobocmf_b200.util.blackbox_mfdgp_fitter import BlackBoxMFDGPFitter, _GraphedConditionedStep
    x, ys, fid = forrester_data()
    N, L, P, T = x.shape[0], 2, 7, 10
    torch.manual_seed(0)
    fitter = BlackBoxMFDGPFitter(L, N, num_epochs_1=0, num_epochs_2=0, device=torch.device(DEV))
    fitter.verbose = False
    fitter.initialize_mfdgp(x, ys["obj1"], fid, "obj1")
    fitter.initialize_mfdgp(x, ys["obj2"], fid, "obj2")
    fitter.initialize_mfdgp(x, ys["con1"], fid, "con1", threshold_constraint=0.1, is_constraint=True)
    g = torch.Generator().manual_seed(3)
    fitter.pareto_set = torch.rand(P, 1, generator=g, dtype=torch.float64).to(DEV)
    fitter.pareto_front = torch.randn(P, 2, generator=g, dtype=torch.float64).to(DEV)
    twin = copy.deepcopy(fitter)
    runs = {}
    for name, ft in (("eager", fitter), ("graph", twin)):
        hobjs, hcons = list(ft.mfdgp_handlers_objs.values()), list(ft.mfdgp_handlers_cons.values())
        params = []
        for h in hobjs + hcons:
            h.mfdgp.fix_variational_hypers_cond(True)
            params += list(h.mfdgp.parameters())
        opt = Adam([{"params": params}], lr=1e-2, capturable=(name == "graph"))
        gstep = _GraphedConditionedStep(ft, hobjs, hcons, opt, static_noise=True) if name == "graph" else None
        gg = torch.Generator().manual_seed(17)
        keys = [("obj", 0), ("obj", 1), ("con", 0)]
        losses = []
        for it in range(5):
            batches, eps = {}, {}
            for key, h in zip(keys, hobjs + hcons):
                perm = torch.randperm(N, generator=gg).to(DEV)
                batches[key] = (h.x[perm], h.y[perm], h.f[perm])
                eps[key] = {w: [None, torch.randn(1, n, generator=gg).double().to(DEV)]
                            for w, n in (("batch", N), ("pareto", P), ("tilde", T))}
            x_tilde = torch.rand(T, 1, generator=gg, dtype=torch.float64).to(DEV)
            if gstep is not None:
                losses.append(float(gstep(batches=batches, x_tilde=x_tilde, eps=eps)))
            else:
                opt.zero_grad()
                loss = ft.conditioned_loss(hobjs, hcons, x_tilde=x_tilde, batches=batches, eps=eps)
                loss.backward()
                opt.step()
                losses.append(float(loss))
        if gstep is not None:
            assert gstep.graphs, "the iteration was not captured"
        runs[name] = (losses, [p.detach().clone() for p in params])
    for a, b in zip(runs["eager"][0], runs["graph"][0]):
        assert abs(a - b) <= 1e-12 * abs(a), (a, b)
    assert runs["eager"][0][-1] != runs["eager"][0][0]
    for a, b in zip(runs["eager"][1], runs["graph"][1]):
        assert relerr(b, a) < 1e-12, relerr(b, a)
