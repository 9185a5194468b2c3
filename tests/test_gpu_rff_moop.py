"""GPU parity of the Pareto-sample row through the C ABI: mobo_rff_eval / mobo_pareto_mask and
the host mirrors on top of them (mobocmf_b200.rff, mobocmf_b200.util.moop) against
  * golden vectors produced by the reference's own code (tests/golden/make_golden_rff_moop.py), and
  * the numpy oracle (oracle/rff_moop_oracle.py) at sizes the goldens do not cover.
Tolerances: function evaluation with given (W, b, theta): 1e-12 relative to max|f| (sums of <= 1500 cos terms);
posterior weights recomputed on the GPU: 1e-6 (A = Phi Phi^T + 1e-6 I has condition number ~1e8-1e9, both sides lose
cond * eps); integer / boolean results (masks, subsets): exact."""
import types

import numpy as np
import pytest
import torch

from tests.test_rff_moop_oracle import golden_chain, load, rel

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def to_sample(chain, d):
    from mobocmf_b200.rff import RFFSample
    dev_chain = []
    for s in chain:
        dev_chain.append({k: (torch.as_tensor(v, dtype=torch.float64, device=DEV) if isinstance(v, np.ndarray) else v)
                          for k, v in s.items()})
    return RFFSample(dev_chain, d, chain[0]["nF"], DEV)


@pytest.mark.parametrize("name", ["rff_d2", "rff_d3"])
@pytest.mark.parametrize("mode", ["posterior", "prior"])
def test_rff_eval_reproduces_reference(name, mode):
    g = load(name)
    d = int(g["d"])
    s = to_sample(golden_chain(g, mode), d)
    f = s.all_layers(g["X"]).cpu().numpy()
    assert rel(f[0], g[mode + "_f0"]) < 1e-12 and rel(f[1], g[mode + "_f1"]) < 1e-12
    assert rel(s(g["X"]), g[mode + "_f1"]) < 1e-12                       # numpy in -> numpy out, top layer
    for i, x in enumerate(g["xg"]):
        assert rel(s(x, gradient=True), g[mode + "_g1"][i]) < 1e-12
    _, gb = s.value_and_grad(g["xg"])
    assert rel(gb.cpu().numpy(), g[mode + "_g1"]) < 1e-12
    s0 = to_sample(golden_chain(g, mode)[:1], d)
    assert rel(s0(g["xg"][0], gradient=True), g[mode + "_g0"][0]) < 1e-12
    import copy
    s2 = copy.deepcopy(s)                       # the fitter holding the samples is deep-copied (copy_uncond)
    assert rel(s2(g["X"]), g[mode + "_f1"]) < 1e-12


@pytest.mark.parametrize("d,L,F,n", [(6, 3, 500, 4099), (1, 2, 37, 33), (8, 4, 64, 257)])
def test_rff_eval_matches_oracle(d, L, F, n):
    from oracle import rff_moop_oracle as R
    rng = np.random.RandomState(10 + d)
    chain = [R.draw_prior_layer0(d, F, rng)] + [R.draw_prior_layer(d, F, rng) for _ in range(L - 1)]
    for s in chain[1:]:
        s["alpha_x1"], s["alpha_x1f"], s["alpha_x2"], s["nu_lin"] = 0.8, 0.9, 0.3, 0.6
    X = rng.uniform(size=(n, d))
    ref = R.eval_chain(chain, X)
    s = to_sample(chain, d)
    f = s.all_layers(X).cpu().numpy()
    for l in range(L):
        assert rel(f[l], ref[l]) < 1e-12, (l, rel(f[l], ref[l]))
    fv, gb = s.value_and_grad(X[:5])
    for i in range(5):
        assert rel(gb[i].cpu().numpy(), R.grad_chain(chain, X[i])[-1]) < 1e-11


def _fake_layer(num_layer, Zx, zf, m, S, theta):
    ns = types.SimpleNamespace
    t = lambda v: torch.as_tensor(np.asarray(v, dtype=np.float64), device=DEV)
    vd = ns(variational_mean=t(m), chol_variational_covar=t(np.linalg.cholesky(S)))
    return ns(num_layer=num_layer, variational_strategy=ns(_variational_distribution=vd), theta=lambda: t(theta),
              _Zx=lambda: t(Zx), _propagated_inducing_column=lambda: t(zf))


@pytest.mark.parametrize("name", ["rff_d2", "rff_d3"])
@pytest.mark.parametrize("mode", ["posterior", "prior"])
def test_sampler_reproduces_reference_stream(name, mode):
    """np.random.seed(s) -> the reference's W, b exactly and its posterior weights / function values to 1e-6."""
    from mobocmf_b200 import rff
    g = load(name)
    d, nF = int(g["d"]), int(g["nF"])
    th0 = np.concatenate([[g["h0_a"]], g["h0_l"]])
    th1 = np.concatenate([[g["h1_a1"], g["h1_v"], g["h1_af"], g["h1_lf"], g["h1_a2"]], g["h1_l1"], g["h1_l2"]])
    L0 = _fake_layer(0, g["Zx"], None, g["m0"], g["S0"], th0)
    L1 = _fake_layer(1, g["Zx"], g["m0"], g["m1"], g["S1"], th1)
    np.random.seed(int(g["seed_np"]))
    s0 = rff.sample_layer(L0, d, None, nF, prior=mode == "prior")
    s1 = rff.sample_layer(L1, d, s0, nF, prior=mode == "prior")
    ref = golden_chain(g, mode)
    for a, b in zip(ref, s1.chain):
        for k, v in a.items():
            if isinstance(v, np.ndarray):
                tol = 1e-6 if (k == "theta" and mode == "posterior") else 1e-15
                assert rel(b[k].cpu().numpy().reshape(v.shape), v) < tol, (k, rel(b[k].cpu().numpy().reshape(v.shape), v))
    assert rel(s1(g["X"]), g[mode + "_f1"]) < 1e-6
    assert rel(s0(g["X"]), g[mode + "_f0"]) < 1e-6


def test_model_function_samples_follow_the_variational_posterior():
    """End to end on the Forrester model: a posterior sample of every layer tracks the variational mean at the
    inducing inputs (q(u) is nearly a point mass at initialisation), and prior samples are O(1) functions."""
    from mobocmf_b200.models.mfdgp import MFDGP
    from tests.helpers import forrester_data
    x, ys, fid = forrester_data()
    torch.manual_seed(0)
    model = MFDGP(x, ys["obj1"], fid, 2)
    model.double().to(DEV)
    np.random.seed(3)
    samples = model.sample_function_from_each_layer()
    assert len(samples) == 2 and samples[1].num_layers == 2
    Z = model.hidden_layer_0._Zx()
    f = samples[1].all_layers(Z)
    # 500 random features approximate the kernel to a few percent: the sample tracks the variational mean (spread 3.3)
    m0 = model.hidden_layer_0.variational_strategy._variational_distribution.variational_mean.detach()
    assert float((f[0] - m0).abs().max()) < 0.3
    m1 = model.hidden_layer_1.variational_strategy._variational_distribution.variational_mean.detach()
    assert float((f[1] - m1).abs().max()) < 0.5
    pri = model.sample_function_from_prior_each_layer(nFeatures=200)
    v = pri[1](np.random.uniform(size=(50, 1)))
    assert v.shape == (50,) and np.all(np.isfinite(v)) and np.abs(v).max() < 50.0


@pytest.mark.parametrize("name", ["moop_k2", "moop_k3"])
def test_pareto_mask_and_summary_reproduce_reference(name):
    from mobocmf_b200.util.moop import MOOP, pareto_mask
    g = load(name)
    pts = torch.as_tensor(g["pts"], device=DEV)
    mask = pareto_mask(pts).cpu().numpy()
    assert np.array_equal(mask, g["mask"])
    d = g["grid"].shape[1]
    mo = MOOP([], [], input_dim=d, grid_size=10, pareto_set_size=7)
    pset, pfront = mo.compute_pareto_front_and_set_summary_y_space(torch.as_tensor(g["pset"], device=DEV),
                                                                   pts[torch.as_tensor(g["mask"], device=DEV)], 7)
    assert np.array_equal(pfront.cpu().numpy(), g["summary_front"])
    assert np.array_equal(pset.cpu().numpy(), g["summary_set"])
    cons = [lambda x: np.sin(4.0 * x[:, 0]) - 0.2, lambda x: x[:, -1] - 0.3]
    fv = np.array([0.1, 0.05] + [0.0] * max(0, d - 2))
    feas = mo.find_feasible_grid(cons, g["grid"], feasible_values=fv)
    assert np.array_equal(feas.cpu().numpy(), g["feasible"])
    hard = [lambda x: -1.0 - x[:, 0], lambda x: -0.5 - x[:, -1] ** 2]
    assert mo.find_feasible_grid(hard, g["grid"], feasible_values=np.zeros(d)) is None
    closest = mo.find_feasible_grid(hard, g["grid"], feasible_values=np.zeros(d), allow_negative_constraints=True)
    assert np.array_equal(closest.cpu().numpy(), g["closest"])


@pytest.mark.parametrize("n,k", [(1, 2), (257, 1), (5000, 4), (20011, 3)])
def test_pareto_mask_matches_bruteforce(n, k):
    from mobocmf_b200.util.moop import pareto_mask
    rng = np.random.RandomState(n + k)
    pts = np.round(rng.normal(size=(n, k)), 2 if n < 6000 else 6)      # coarse rounding -> ties and duplicates
    mask = pareto_mask(torch.as_tensor(pts, device=DEV)).cpu().numpy()
    P = torch.as_tensor(pts, device=DEV)
    dom = torch.zeros(n, dtype=torch.bool, device=DEV)
    idx = torch.arange(n, device=DEV)
    for j0 in range(0, n, 512):
        blk = P[j0:j0 + 512]
        le = (P[None, :, :] <= blk[:, None, :]).all(-1)
        lt = (P[None, :, :] < blk[:, None, :]).any(-1)
        first = idx[None, :] < idx[j0:j0 + 512, None]
        dom[j0:j0 + 512] = (le & (lt | first)).any(1)
    assert np.array_equal(mask, (~dom).cpu().numpy())


def test_full_search_reproduces_reference():
    """MOOP.compute_pareto_solution_from_samples on closed-form samples: same numpy seed -> same grid; the GPU cull and
    summary select the same Pareto set as the reference run stored in moop_full_d2.npz."""
    from mobocmf_b200.util.moop import MOOP
    from tests.golden.make_golden_rff_moop import closed_form_problem
    g = load("moop_full_d2")
    d = int(g["d"])
    objs, cons = closed_form_problem(d)
    np.random.seed(int(g["seed_np"]))
    mo = MOOP(objs, cons, input_dim=d, grid_size=int(g["grid_size"]), pareto_set_size=int(g["pareto_set_size"]),
              feasible_values=np.zeros(d))
    pset, pfront, _, _ = mo.compute_pareto_solution_from_samples(torch.as_tensor(g["inputs"]))
    assert pset.shape == g["pareto_set"].shape
    assert np.abs(pset.numpy() - g["pareto_set"]).max() < 1e-9
    assert np.abs(pfront.numpy() - g["pareto_front"]).max() < 1e-9


def test_fitter_samples_a_pareto_solution():
    """BlackBoxMFDGPFitter.sample_and_store_pareto_solution on two objectives + one constraint (Forrester-sized)."""
    from mobocmf_b200.util.blackbox_mfdgp_fitter import BlackBoxMFDGPFitter
    from tests.helpers import forrester_data
    x, ys, fid = forrester_data()
    torch.manual_seed(0)
    np.random.seed(0)
    fitter = BlackBoxMFDGPFitter(2, batch_size=16, num_epochs_1=1, num_epochs_2=1, pareto_set_size=10,
                                 opt_grid_size=200)
    fitter.initialize_mfdgp(x, ys["obj1"], fid, "obj1")
    fitter.initialize_mfdgp(x, -ys["obj1"] + 0.3 * x, fid, "obj2")
    fitter.initialize_mfdgp(x, ys["obj1"] * 0 + 1.0 - x, fid, "con1", is_constraint=True)
    pset, pfront, so, sc = fitter.sample_and_store_pareto_solution()
    assert pset.shape[1] == 1 and pfront.shape[1] == 2 and 1 <= pset.shape[0] <= 10
    assert torch.all((pset >= 0) & (pset <= 1))
    # the stored front is the objective samples evaluated at the stored set, and no stored point dominates another
    vals = torch.stack([torch.as_tensor(s(pset.numpy())) for s in so], dim=1)
    assert float((vals - pfront).abs().max()) < 1e-10
    for i in range(pset.shape[0]):
        assert not bool(((pfront <= pfront[i]).all(1) & (pfront < pfront[i]).any(1)).any())
    assert float(torch.as_tensor(sc[0](pset.numpy())).min()) >= -1e-9
