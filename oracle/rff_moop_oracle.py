"""CPU (numpy / scipy, fp64) restatement of the reference's Pareto-sample generation: random-Fourier-feature function
samples of every MFDGP layer and the multi-objective grid search over them.

TEST INFRASTRUCTURE ONLY: imported by tests/ (and nothing else); the product path is mobocmf_b200/rff.py,
mobocmf_b200/util/moop.py over csrc/rff.cu.  PINNED: tests/golden/{rff_d2,rff_d3,moop_k2,moop_k3}.npz were produced by
running the reference's own code (tests/golden/make_golden_rff_moop.py) and tests/test_rff_moop_oracle.py checks this
file against them.

Reference lines restated (relative to /root/reference/mobocmf/):
  layers/mfdgp_hidden_layer.py:288-293   _phi_rbf                        -> phi_rbf
  layers/mfdgp_hidden_layer.py:295-309   _chol2inv, _rff_sample_posterior_weights -> posterior_weights
  layers/mfdgp_hidden_layer.py:311-338   _sample_from_posterior_layer0   -> draw_posterior_layer0
  layers/mfdgp_hidden_layer.py:340-362   _sample_from_prior_layer0       -> draw_prior_layer0
  layers/mfdgp_hidden_layer.py:364-444   _sample_from_posterior          -> draw_posterior_layer, eval_chain, grad_chain
  layers/mfdgp_hidden_layer.py:446-514   _sample_from_prior              -> draw_prior_layer
  models/mfdgp.py:264-288                sample_function_from(_prior)_each_layer -> the chains below
  util/moop.py:38-70                     find_feasible_grid              -> feasible_grid
  util/moop.py:141-185                   compute_pareto_front / obtain_indices_pareto -> pareto_mask
  util/moop.py:187-219                   compute_pareto_front_and_set_summary_y_space -> summary_subset
A function sample is a dict (one per layer); a chain is the list of them from the lowest fidelity up.  All random
numbers come from ``rng`` (default: numpy's global generator, drawn in the reference's order, so ``np.random.seed(s)``
reproduces the reference's samples)."""
import numpy as np
import scipy.linalg as spla


# ------------------------------------------------------------------------------------------------
# random Fourier features
# ------------------------------------------------------------------------------------------------
def phi_rbf(x, W, b, alpha, nF, gradient=False):
    """Features of an RBF kernel with amplitude alpha: sqrt(2 alpha / nF) cos(W x^T + b), shape (nF, n); with
    gradient=True (single point only) the (nF, d) matrix of d feature / d x."""
    scale = np.sqrt(2.0 * alpha / nF)
    arg = W @ x.T + b
    if gradient:
        return -scale * np.sin(arg) * W
    return scale * np.cos(arg)


def posterior_weights(y, S, Phi, randomness, sigma2=1e-6):
    """theta ~ N(A^-1 Phi y, sigma2 A^-1 + A^-1 Phi S Phi^T A^-1), A = Phi Phi^T + sigma2 I; `randomness` are the
    standard normals the reference draws first (size = number of features)."""
    nfeat = Phi.shape[0]
    A = Phi @ Phi.T + sigma2 * np.eye(nfeat)
    cA = spla.cholesky(A)                                    # upper factor
    A_inv = spla.cho_solve((cA, False), np.eye(nfeat))
    mean = spla.cho_solve((cA, False), Phi @ y)
    extra = (A_inv @ Phi) @ S @ (Phi.T @ A_inv)
    return mean + (randomness @ spla.cholesky(sigma2 * A_inv + extra, lower=False)).T


def _rng(rng):
    return np.random if rng is None else rng


def draw_posterior_layer0(Z, m, S, lengthscale, alpha, nF=500, rng=None):
    """Z (M, d) inducing inputs, m (M,) variational mean, S (M, M) variational covariance."""
    r = _rng(rng)
    d = Z.shape[1]
    W = r.normal(size=(nF, d)) / lengthscale
    b = r.uniform(low=0.0, high=2 * np.pi, size=(nF, 1))
    Phi = phi_rbf(Z, W, b, alpha, nF)
    randomness = r.normal(loc=0.0, scale=1.0, size=Phi.shape[0])
    theta = posterior_weights(m, S, Phi, randomness)
    return dict(kind=0, nF=nF, W=W, b=b, theta=theta, alpha=float(alpha))


def draw_prior_layer0(d, nF=500, rng=None):
    r = _rng(rng)
    W = r.normal(size=(nF, d)) / (0.25 * d)
    b = r.uniform(low=0.0, high=2 * np.pi, size=(nF, 1))
    theta = r.normal(loc=0.0, scale=1.0, size=nF)
    return dict(kind=0, nF=nF, W=W, b=b, theta=theta, alpha=1.0)


def _draw_upper_features(d, nF, l1, lf, l2, r):
    W_x1 = r.normal(size=(nF, d)) / l1
    W_f = r.normal(size=nF) / lf
    W_x2 = r.normal(size=(nF, d)) / l2
    b_x1 = r.uniform(low=0.0, high=2 * np.pi, size=(nF, 1))
    b_x2 = r.uniform(low=0.0, high=2 * np.pi, size=(nF, 1))
    return W_x1, W_f, W_x2, b_x1, b_x2


def draw_posterior_layer(Zxf, m, S, l1, lf, l2, a1, af, a2, nu, nF=500, rng=None):
    """Zxf (M, d + 1): inducing inputs with the propagated column last."""
    r = _rng(rng)
    d = Zxf.shape[1] - 1
    Zx, zf = Zxf[:, :d], Zxf[:, d]
    W_x1, W_f, W_x2, b_x1, b_x2 = _draw_upper_features(d, nF, l1, lf, l2, r)
    a1f = a1 * af
    W_x1f = np.concatenate([W_x1, W_f[:, None]], axis=1)
    Phi = np.concatenate([phi_rbf(Zx, W_x1, b_x1, a1, nF) * zf * np.sqrt(nu),
                          phi_rbf(Zxf, W_x1f, b_x1, a1f, nF),
                          phi_rbf(Zx, W_x2, b_x2, a2, nF)])
    randomness = r.normal(loc=0.0, scale=1.0, size=Phi.shape[0])
    theta = posterior_weights(m, S, Phi, randomness)
    return dict(kind=1, nF=nF, W_x1=W_x1, W_f=W_f, W_x2=W_x2, b_x1=b_x1, b_x2=b_x2, theta=theta, alpha_x1=float(a1),
                alpha_x1f=float(a1f), alpha_x2=float(a2), nu_lin=float(nu))


def draw_prior_layer(d, nF=500, rng=None):
    r = _rng(rng)
    W_x1, W_f, W_x2, b_x1, b_x2 = _draw_upper_features(d, nF, 10 * 0.25 * d, 1.0, 0.25 * d, r)
    theta = r.normal(loc=0.0, scale=1.0, size=3 * nF)
    return dict(kind=1, nF=nF, W_x1=W_x1, W_f=W_f, W_x2=W_x2, b_x1=b_x1, b_x2=b_x2, theta=theta, alpha_x1=1.0,
                alpha_x1f=1.0, alpha_x2=0.01, nu_lin=1.0)


def eval_chain(chain, x):
    """Values of every layer's function sample at the rows of x (n, d): list of (n,) arrays, lowest fidelity first."""
    x = np.atleast_2d(x)
    out, f = [], None
    for s in chain:
        nF = s["nF"]
        if s["kind"] == 0:
            f = s["theta"] @ phi_rbf(x, s["W"], s["b"], s["alpha"], nF)
        else:
            xf = np.concatenate([x, f[:, None]], axis=1)
            W_x1f = np.concatenate([s["W_x1"], s["W_f"][:, None]], axis=1)
            feats = np.concatenate([phi_rbf(x, s["W_x1"], s["b_x1"], s["alpha_x1"], nF) * f * np.sqrt(s["nu_lin"]),
                                    phi_rbf(xf, W_x1f, s["b_x1"], s["alpha_x1f"], nF),
                                    phi_rbf(x, s["W_x2"], s["b_x2"], s["alpha_x2"], nF)])
            f = s["theta"] @ feats
        out.append(f)
    return out


def grad_chain(chain, x):
    """d f_l / d x at ONE point x (d,): list of (d,) arrays."""
    x = np.atleast_2d(x)
    assert x.shape[0] == 1
    vals = eval_chain(chain, x)
    out, df = [], None
    for l, s in enumerate(chain):
        nF = s["nF"]
        if s["kind"] == 0:
            df = s["theta"] @ phi_rbf(x, s["W"], s["b"], s["alpha"], nF, gradient=True)
        else:
            f = vals[l - 1]
            xf = np.concatenate([x, f[:, None]], axis=1)
            W_x1f = np.concatenate([s["W_x1"], s["W_f"][:, None]], axis=1)
            feat_x1 = phi_rbf(x, s["W_x1"], s["b_x1"], s["alpha_x1"], nF)
            dxf_dx = np.concatenate([np.eye(x.shape[1]), df[:, None]], axis=1)              # (d, d + 1)
            d_x1 = phi_rbf(x, s["W_x1"], s["b_x1"], s["alpha_x1"], nF, gradient=True)
            d_x1f = phi_rbf(xf, W_x1f, s["b_x1"], s["alpha_x1f"], nF, gradient=True) @ dxf_dx.T
            d_x2 = phi_rbf(x, s["W_x2"], s["b_x2"], s["alpha_x2"], nF, gradient=True)
            feats = np.concatenate([(d_x1 * f + df * feat_x1) * np.sqrt(s["nu_lin"]), d_x1f, d_x2])
            df = s["theta"] @ feats
        out.append(np.asarray(df).reshape(-1))
    return out


# ------------------------------------------------------------------------------------------------
# multi-objective grid search
# ------------------------------------------------------------------------------------------------
def feasible_grid(con_values, grid, feasible_values, allow_negative_constraints=False):
    """con_values: list of (n,) constraint samples on the grid; keeps rows with every c_i >= feasible_values[i].  With
    allow_negative_constraints and an empty feasible set: the row(s) whose summed violation is smallest."""
    ok = np.ones(grid.shape[0], dtype=bool)
    for i, c in enumerate(con_values):
        ok &= c >= feasible_values[i]
    if ok.any():
        return grid[ok]
    if not allow_negative_constraints:
        return None
    viol = np.zeros(grid.shape[0])
    for i, c in enumerate(con_values):
        viol += np.minimum(c - feasible_values[i], 0.0)
    return grid[viol == np.max(viol[viol != 0])]


def pareto_mask(pts):
    """Boolean mask of the points of pts (n, k) that survive the reference's cull: visited in the order of their
    standardised coordinate sum, a point removes every other point that is nowhere strictly smaller than it."""
    n = pts.shape[0]
    order = np.argsort(((pts - pts.mean(axis=0)) / (pts.std(axis=0) + 1e-7)).sum(axis=1))
    alive = list(order)
    pos = 0
    while pos < len(alive):
        p = pts[alive[pos]]
        keep = [q for j, q in enumerate(alive) if j == pos or np.any(pts[q] < p)]
        pos = keep.index(alive[pos]) + 1
        alive = keep
    mask = np.zeros(n, dtype=bool)
    mask[alive] = True
    return mask


def summary_subset(front, size):
    """Indices of the `size` front points the reference keeps: the best of each objective, then greedily the point
    farthest (in objective space) from those already chosen."""
    n, k = front.shape
    if n <= size:
        return np.arange(n)
    dist = np.sqrt(((front[:, None, :] - front[None, :, :]) ** 2).sum(-1))
    chosen = np.zeros(size, dtype=np.int64)
    for i in range(k):
        chosen[i] = np.argmin(front[:, i])
    for c in range(k, size):
        chosen[c] = np.argmax(np.min(dist[chosen[:c], :], axis=0))
    return chosen
