"""Pins oracle/rff_moop_oracle.py against golden vectors produced by the reference's own code
(tests/golden/make_golden_rff_moop.py, run in the build container where /root/reference exists)."""
import os

import numpy as np
import pytest

from oracle import rff_moop_oracle as R

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    z = np.load(os.path.join(HERE, name + ".npz"))
    return {k: z[k] for k in z.files}


def golden_chain(g, mode):
    """The reference's own draws (W, b, theta read back from its closures) as an oracle chain."""
    nF = int(g["nF"])
    p = mode + "_"
    s0 = dict(kind=0, nF=nF, W=g[p + "W0"], b=g[p + "b0"], theta=g[p + "theta0"], alpha=float(g[p + "alpha0"]))
    Wx1f = g[p + "1_W_x1f"]
    s1 = dict(kind=1, nF=nF, W_x1=g[p + "1_W_x1"], W_f=Wx1f[:, -1], W_x2=g[p + "1_W_x2"], b_x1=g[p + "1_b_x1"],
              b_x2=g[p + "1_b_x2"], theta=g[p + "1_theta"], alpha_x1=float(g[p + "1_alpha_x1"]),
              alpha_x1f=float(g[p + "1_alpha_x1f"]), alpha_x2=float(g[p + "1_alpha_x2"]), nu_lin=float(g[p + "1_nu_lin"]))
    return [s0, s1]


def oracle_chain(g, mode):
    """The oracle's draws from the same numpy seed and the same model state."""
    d, nF = int(g["d"]), int(g["nF"])
    np.random.seed(int(g["seed_np"]))
    if mode == "prior":
        return [R.draw_prior_layer0(d, nF), R.draw_prior_layer(d, nF)]
    s0 = R.draw_posterior_layer0(g["Zx"], g["m0"], g["S0"], g["h0_l"], float(g["h0_a"]), nF)
    Zxf = np.concatenate([g["Zx"], g["m0"][:, None]], axis=1)
    s1 = R.draw_posterior_layer(Zxf, g["m1"], g["S1"], g["h1_l1"], float(g["h1_lf"]), g["h1_l2"], float(g["h1_a1"]),
                                float(g["h1_af"]), float(g["h1_a2"]), float(g["h1_v"]), nF)
    return [s0, s1]


def rel(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(1e-300, np.max(np.abs(b))))


@pytest.mark.parametrize("name", ["rff_d2", "rff_d3"])
@pytest.mark.parametrize("mode", ["posterior", "prior"])
def test_function_evaluation_matches_reference(name, mode):
    g = load(name)
    chain = golden_chain(g, mode)
    f = R.eval_chain(chain, g["X"])
    assert rel(f[0], g[mode + "_f0"]) < 1e-13 and rel(f[1], g[mode + "_f1"]) < 1e-13
    for i, x in enumerate(g["xg"]):
        df = R.grad_chain(chain, x)
        assert rel(df[0], g[mode + "_g0"][i]) < 1e-13 and rel(df[1], g[mode + "_g1"][i]) < 1e-13


@pytest.mark.parametrize("name", ["rff_d2", "rff_d3"])
@pytest.mark.parametrize("mode", ["posterior", "prior"])
def test_draws_match_reference_stream(name, mode):
    """Same numpy seed -> same W, b and (through the posterior-weight algebra) the same theta as the reference."""
    g = load(name)
    ref, mine = golden_chain(g, mode), oracle_chain(g, mode)
    for a, b in zip(ref, mine):
        for k in a:
            if isinstance(a[k], np.ndarray):
                tol = 1e-9 if k == "theta" and mode == "posterior" else 1e-15
                assert rel(b[k], a[k]) < tol, (k, rel(b[k], a[k]))
            else:
                assert a[k] == pytest.approx(b[k], rel=1e-15)
    f = R.eval_chain(mine, g["X"])
    assert rel(f[1], g[mode + "_f1"]) < 1e-8


def test_gradient_is_the_derivative():
    g = load("rff_d3")
    chain = golden_chain(g, "posterior")
    x = g["xg"][0]
    df = R.grad_chain(chain, x)[1]
    h = 1e-6
    for c in range(len(x)):
        e = np.zeros_like(x); e[c] = h
        num = (R.eval_chain(chain, x + e)[1] - R.eval_chain(chain, x - e)[1]) / (2 * h)
        assert abs(num[0] - df[c]) < 1e-6 * max(1.0, abs(df[c]))


@pytest.mark.parametrize("name", ["moop_k2", "moop_k3"])
def test_moop_matches_reference(name):
    g = load(name)
    mask = R.pareto_mask(g["pts"])
    assert np.array_equal(mask, g["mask"])
    front = g["pts"][g["mask"]]
    idx = R.summary_subset(front, 7)
    assert np.array_equal(front[idx], g["summary_front"]) and np.array_equal(g["pset"][idx], g["summary_set"])
    grid = g["grid"]
    d = grid.shape[1]
    cons = [np.sin(4.0 * grid[:, 0]) - 0.2, grid[:, -1] - 0.3]
    fv = np.array([0.1, 0.05] + [0.0] * max(0, d - 2))
    assert np.array_equal(R.feasible_grid(cons, grid, fv), g["feasible"])
    hard = [-1.0 - grid[:, 0], -0.5 - grid[:, -1] ** 2]
    assert R.feasible_grid(hard, grid, np.zeros(d)) is None
    assert np.array_equal(R.feasible_grid(hard, grid, np.zeros(d), allow_negative_constraints=True), g["closest"])


def test_pareto_mask_is_the_nondominated_set():
    rng = np.random.RandomState(0)
    pts = rng.normal(size=(300, 3))
    mask = R.pareto_mask(pts)
    for j in range(300):
        dominated = np.any(np.all(pts <= pts[j], axis=1) & np.any(pts < pts[j], axis=1))
        assert mask[j] == (not dominated)
