"""Extended-precision TRUTH for the MFDGP hot path: the functions of ``oracle/mfdgp_oracle.py`` restated over
``oracle/ld_autodiff.py`` (numpy longdouble, 64-bit mantissa, reverse-mode gradients).

TEST INFRASTRUCTURE ONLY.  Purpose ("make gradients evidence"): with the reference's default
initialisation cond(K_zz + jitter I) is 1e7 - 1e8 and two correct fp64 implementations of the same formulas differ by
~cond * eps, i.e. 1e-8 in values and 1e-3 in some gradients.  A fixed loose bar cannot tell such rounding from a
missing term.  Here both the fp64 oracle and the CUDA path are compared with a result whose own error is ~2000x
smaller, and the tests assert ``|cuda - truth| <= c * |oracle - truth|`` (c = 10) or 1e-10, whichever is larger, for
the loss and every gradient: the CUDA path must be as close to the exact answer as the reference-shaped fp64
formulation is.

The formulas follow the same reference lines as the oracle (see its header) with two differences that only matter
in finite precision: squared distances are sums of squared differences (the oracle reproduces upstream's centred
quadratic expansion) and the predictive terms use triangular solves throughout.  ``tests/test_truth.py`` pins this
module against the fp64 oracle (values and every gradient, well-conditioned inputs: agreement ~1e-13).
"""
import math

import numpy as np

from . import ld_autodiff as A
from .ld_autodiff import LD

JITTER = 1e-6
MIN_VARIANCE = 1e-10
NOISE_LOWER = 1e-8


def _key_layer(l):
    return "hidden_layer_%d." % l


def leaves(sd, names):
    """sd (torch fp64 state dict, GPyTorch naming) -> dict of tape nodes; ``names`` become differentiable leaves."""
    out = {}
    for k, v in sd.items():
        if not v.dtype.is_floating_point:
            continue
        out[k] = A.leaf(v) if k in names else A.const(v)
    return out


def layer_hypers(sd, l):
    p = _key_layer(l) + "covar_module."
    sp = A.softplus
    if l == 0:
        return {"a": sp(sd[p + "raw_outputscale"]), "ls": sp(sd[p + "base_kernel.raw_lengthscale"])}
    return {"a1": sp(sd[p + "kernels.0.kernels.0.raw_outputscale"]),
            "ls1": sp(sd[p + "kernels.0.kernels.0.base_kernel.raw_lengthscale"]),
            "vlin": A.reshape(sp(sd[p + "kernels.0.kernels.1.kernels.0.raw_variance"]), ()),
            "af": sp(sd[p + "kernels.0.kernels.1.kernels.1.raw_outputscale"]),
            "lsf": sp(sd[p + "kernels.0.kernels.1.kernels.1.base_kernel.raw_lengthscale"]),
            "a2": sp(sd[p + "kernels.1.raw_outputscale"]),
            "ls2": sp(sd[p + "kernels.1.base_kernel.raw_lengthscale"])}


def rbf(x1, x2, ls):
    """exp(-1/2 sum_c ((x1_c - x2_c) / l_c)^2), (n1, n2).  ls: (1, d)."""
    a = A.reshape(x1 / ls, (x1.shape[0], 1, x1.shape[1]))
    b = A.reshape(x2 / ls, (1, x2.shape[0], x2.shape[1]))
    return A.exp(A.sum_(A.square(a - b), axis=2) * -0.5)


def layer_kernel(sd, l, Xa, Xb):
    h = layer_hypers(sd, l)
    if l == 0:
        return h["a"] * rbf(Xa, Xb, h["ls"])
    d = Xa.shape[1] - 1
    xa, xb, fa, fb = Xa[:, :d], Xb[:, :d], Xa[:, d:], Xb[:, d:]
    k_x1 = h["a1"] * rbf(xa, xb, h["ls1"])
    k_lin = h["vlin"] * (fa @ fb.T)
    k_f = h["af"] * rbf(fa, fb, h["lsf"])
    k_x2 = h["a2"] * rbf(xa, xb, h["ls2"])
    return k_x1 * (k_lin + k_f) + k_x2


def layer_kernel_diag(sd, l, Xa):
    h = layer_hypers(sd, l)
    one = np.ones(Xa.shape[0], dtype=LD)
    if l == 0:
        return h["a"] * one
    f = Xa[:, Xa.shape[1] - 1]
    return h["a1"] * (h["vlin"] * f * f + h["af"]) + h["a2"] * one


def variational_q(sd, l):
    p = _key_layer(l) + "variational_strategy._variational_distribution."
    return sd[p + "variational_mean"], A.tril(sd[p + "chol_variational_covar"])


def layer_inducing_points(sd, l):
    """Z_l = [Z[:, :d], previous_layer(Z[:, :d]).mean] (layers/mfdgp_hidden_layer.py:542-559).  With shared inducing
    inputs quirk Q4 makes the previous layer return its variational mean exactly; otherwise (only-highest-fidelity
    models keep per-layer inducing inputs) layer 0 is evaluated at them."""
    Zo = sd[_key_layer(l) + "variational_strategy.inducing_points"]
    if l == 0:
        return Zo
    d = Zo.shape[1] - 1
    zx = Zo[:, :d]
    zprev = sd[_key_layer(l - 1) + "variational_strategy.inducing_points"]
    if zprev.v.shape[0] == zx.v.shape[0] and np.array_equal(zx.v, zprev.v[:, :d]):
        mean_prev = variational_q(sd, l - 1)[0]
    else:
        assert l == 1, "non-shared inducing inputs above layer 1 crash in the reference too (F3)"
        mean_prev, _ = layer_q(sd, 0, zx, True)
    return A.cat([zx, A.reshape(mean_prev, (-1, 1))], axis=1)


def prior_cholesky(sd, l, Z, jitter=JITTER):
    M = Z.shape[0]
    return A.cholesky(layer_kernel(sd, l, Z, Z) + LD(jitter) * np.eye(M, dtype=LD))


def layer_q(sd, l, X, training=True, jitter=JITTER):
    """(mean, raw variance) of q(f_l(X)); same branches as the oracle's ``layer_q``."""
    Z = layer_inducing_points(sd, l)
    m, Lq = variational_q(sd, l)
    Lp = prior_cholesky(sd, l, Z, jitter)
    Kzx = layer_kernel(sd, l, Z, X)
    kxx = layer_kernel_diag(sd, l, X)
    T = A.solve_lower(Lp, Kzx)                       # L^-1 K_zx
    Am = A.solve_lower_t(Lp, T)                      # P^-1 K_zx
    mean = A.reshape(A.reshape(m, (1, -1)) @ Am, (-1,))
    q = A.sum_(A.square(T), axis=0)
    data_var = A.clamp_min(kxx - q, 0.0) if training else kxx - q
    var = A.sum_(A.square(Lq.T @ Am), axis=0) + data_var
    return mean, var


def kl_layer(sd, l, jitter=JITTER):
    Z = layer_inducing_points(sd, l)
    Lp = prior_cholesky(sd, l, Z, jitter)
    m, Lq = variational_q(sd, l)
    M = m.shape[0]
    R = A.solve_lower(Lp, A.cat([A.reshape(m, (-1, 1)), Lq], axis=1))
    idx = np.arange(M)
    logdet_p = A.sum_(A.log(A.square(Lp[idx, idx])))
    logdet_q = A.sum_(A.log(A.square(Lq[idx, idx])))
    return 0.5 * (logdet_p - logdet_q + A.sum_(A.square(R)) - float(M))


def read_variance(v):
    return A.clamp_min(v, MIN_VARIANCE)


def noise_value(sd, l, upper, lower):
    raw = sd["hidden_layer_likelihood_%d.noise_covar.raw_noise" % l]
    lo = lower[l] if isinstance(lower, (list, tuple)) else lower
    return A.reshape(LD(lo) + (LD(upper[l]) - LD(lo)) * A.sigmoid(raw), ())


def expected_log_prob(target, mean, var, noise):
    return ((A.square(target - mean) + var) / noise + A.log(noise) + LD(math.log(2 * math.pi))) * -0.5


def elbo_step_loss_tiled(sd, L, noise_upper, x, y, fid, eps, num_data, S, noise_lower=NOISE_LOWER, jitter=JITTER,
                         x_node=None, only_hf=False):
    """-ELBO of the S-sample tiled step (the oracle's ``elbo_step_loss_tiled``; S = 1 with (1, B) normals is the
    reference's own step).  x, y, fid, eps: torch tensors / arrays (constants); returns (loss node, kl node)."""
    X = A.const(x) if x_node is None else x_node
    B = X.shape[0]
    yv = np.asarray(y.detach().cpu().numpy() if hasattr(y, "detach") else y, dtype=LD).reshape(-1)
    fv = np.asarray(fid.detach().cpu().numpy() if hasattr(fid, "detach") else fid).reshape(-1)
    mean, var = layer_q(sd, 0, X, True, jitter)
    outs = [(mean, var)]
    x_tile = A.repeat_interleave(X, S, 0)
    pm, pv = A.repeat_interleave(mean, S, 0), A.repeat_interleave(var, S, 0)
    for l in range(1, L):
        if only_hf:                                      # models/mfdgp.py:189-190: the passed mean is zeroed
            f = A.reshape(pm * 0.0, (-1, 1))
        else:
            e = np.asarray(eps[l].detach().cpu().numpy() if hasattr(eps[l], "detach") else eps[l],
                           dtype=LD).reshape(-1)
            f = A.reshape(e * A.sqrt(read_variance(pv)) + pm, (-1, 1))
        pm, pv = layer_q(sd, l, A.cat([x_tile, f], axis=1), True, jitter)
        outs.append((pm, pv))
    data = A.const(0.0)
    for l in range(L):
        mask = (fv == l)
        if mask.sum() == 0:
            continue
        noise = noise_value(sd, l, noise_upper, noise_lower)
        m_, v_ = outs[l]
        if l == 0:
            ell = expected_log_prob(yv, m_, read_variance(v_), noise)
        else:
            ell = expected_log_prob(np.repeat(yv, S), m_, read_variance(v_), noise)
            ell = A.sum_(A.reshape(ell, (B, S)), axis=1) / LD(S)
        data = data + A.sum_(ell * mask.astype(LD))
    kl = A.const(0.0)
    for l in range(L):
        kl = kl + kl_layer(sd, l, jitter)
    scale = LD(B) / LD(num_data)
    return -(data - kl * scale), kl * scale


def predict_for_acquisition(sd, L, noise_upper, samples, X, fidelity, noise_lower=NOISE_LOWER, jitter=JITTER):
    """(mean, variance) per candidate: eval-mode chain over S fixed normals per layer + moment matching
    (models/mfdgp.py:237-262).  X: tape node (n, d) so that d/dX is available."""
    n = X.shape[0]
    S = samples[1].shape[0] if L > 1 else 1
    x_tile = A.repeat_interleave(X, S, 0)
    mean, var = layer_q(sd, 0, x_tile, False, jitter)
    for l in range(1, fidelity + 1):
        s = np.asarray(samples[l].detach().cpu().numpy() if hasattr(samples[l], "detach") else samples[l],
                       dtype=LD).reshape(-1)
        rep = np.tile(s, n)                                      # row i*S+s uses samples[l][s]
        f = A.reshape(mean + A.sqrt(read_variance(var)) * rep, (-1, 1))
        mean, var = layer_q(sd, l, A.cat([x_tile, f], axis=1), False, jitter)
    noise = noise_value(sd, fidelity, noise_upper, noise_lower)
    vt = read_variance(var + noise)
    mus = A.sum_(A.reshape(mean, (n, S)), axis=1) / LD(S)
    second = A.sum_(A.reshape(vt + A.square(mean), (n, S)), axis=1) / LD(S)
    return mus, second - A.square(mus)


def jes(vu, vc):
    return A.clamp_min(A.log(vu) - A.log(vc), 0.0) * 0.5


# ---- convenience drivers used by the tests --------------------------------------------------------------------
def elbo_step_truth(sd_torch, names, L, noise_upper, x, y, fid, eps, num_data, S, noise_lower=NOISE_LOWER,
                    only_hf=False):
    """(loss, kl, {name: gradient}) as longdouble arrays."""
    sd = leaves(sd_torch, set(names))
    loss, kl = elbo_step_loss_tiled(sd, L, noise_upper, x, y, fid, eps, num_data, S, noise_lower=noise_lower,
                                    only_hf=only_hf)
    A.backward(loss)
    grads = {n: (sd[n].g if sd[n].g is not None else np.zeros(sd[n].v.shape, dtype=LD)) for n in names}
    return loss.v, kl.v, grads


class covariance_rounding(object):
    """Context manager: every covariance matrix the truth evaluates inside it is multiplied element by element by
    (1 + eps64 * xi), xi uniform in [-1, 1] (seeded; K(Z_l, Z_l) symmetrically and with the same draw wherever it is
    used).  The change of a result under it is that result's sensitivity to ROUNDING THE COVARIANCE ENTRIES TO fp64 -
    the forward error an fp64 program that forms K may show however carefully it then solves with it.  Unlike
    eps * cond(K_zz) it is specific to each quantity: a gradient that is a small difference of large terms (layer 0's
    outputscale at the reference's default initialisation: every term depends on it only through jitter / a) has a large
    one."""

    def __init__(self, seed, scale=2.220446049250313e-16):
        self.rng, self.scale, self.zz = np.random.default_rng(seed), scale, {}

    def __enter__(self):
        global layer_kernel
        self.orig = layer_kernel

        def noisy(sd, l, Xa, Xb):
            K = self.orig(sd, l, Xa, Xb)
            if Xa is Xb:
                if l not in self.zz:
                    xi = self.rng.uniform(-1.0, 1.0, size=K.shape)
                    self.zz[l] = np.triu(xi) + np.triu(xi, 1).T
                xi = self.zz[l]
            else:
                xi = self.rng.uniform(-1.0, 1.0, size=K.shape)
            return K * (LD(1) + LD(self.scale) * xi.astype(LD))
        layer_kernel = noisy
        return self

    def __exit__(self, *exc):
        global layer_kernel
        layer_kernel = self.orig
        return False


def elbo_step_rounding_sensitivity(sd_torch, names, L, noise_upper, x, y, fid, eps, num_data, S, truth, draws=6,
                                   noise_lower=NOISE_LOWER, only_hf=False):
    """{name or "loss": max over ``draws`` of max |truth(perturbed K) - truth| / max |truth|}; truth = (loss, grads) as
    returned by ``elbo_step_truth`` on the same arguments."""
    loss_t, grads_t = truth
    out = {n: 0.0 for n in list(names) + ["loss"]}
    for k in range(draws):
        with covariance_rounding(1000 + k):
            loss_p, _, grads_p = elbo_step_truth(sd_torch, names, L, noise_upper, x, y, fid, eps, num_data, S,
                                                 noise_lower=noise_lower, only_hf=only_hf)
        out["loss"] = max(out["loss"], float(abs(loss_p - loss_t) / abs(loss_t)))
        for n in names:
            den = np.max(np.abs(grads_t[n]))
            out[n] = max(out[n], float(np.max(np.abs(grads_p[n] - grads_t[n])) / (den if den > 0 else LD(1))))
    return out


def jes_truth(mod_u, mod_c, X, fidelity):
    """(jes values (n,), d sum(jes) / dX (n, d)) for one black box; mod_* as in the oracle's ``jes_mfdgp``."""
    Xn = A.leaf(X.reshape(-1, X.shape[-1]))
    outs = []
    for mod in (mod_u, mod_c):
        sd = leaves(mod["sd"], set())
        _, v = predict_for_acquisition(sd, mod["num_layers"], mod["noise_upper"], mod["samples"], Xn, fidelity,
                                       noise_lower=mod.get("noise_lower", NOISE_LOWER))
        outs.append(v)
    val = jes(outs[0], outs[1])
    A.backward(A.sum_(val))
    return val.v, Xn.g


def err_vs(x, truth):
    """max |x - truth| / max |truth| with x an fp64 torch tensor / array and truth a longdouble array."""
    xv = np.asarray(x.detach().cpu().numpy() if hasattr(x, "detach") else x, dtype=LD).reshape(np.shape(truth))
    den = np.max(np.abs(truth))
    return float(np.max(np.abs(xv - truth)) / (den if den > 0 else LD(1)))
