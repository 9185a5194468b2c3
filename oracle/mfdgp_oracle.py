"""CPU fp64 oracle for the MOBOCMF hot path (MFDGP ELBO step + JES acquisition evaluation).

TEST INFRASTRUCTURE ONLY.  Nothing under ``mobocmf_b200/`` may import this module; only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` do, and
there only as the checker / the reported CPU baseline, never as the product path.

PARITY UNPINNED.  The arithmetic of the reference lives in GPyTorch / linear_operator / BoTorch, which are
not vendored in ``/root/reference``, carry no pinned version (``setup.py:1-17``; ``README.md:14-16`` says only
"botorch >= 0.9.0", whose own pins are gpytorch 1.11 / linear_operator 0.5.1) and are not installable here.
The reference ships no tests, golden vectors or stored outputs.  This file therefore restates,
line by line, the reference's own files plus the published upstream algorithms they call:

* ``mobocmf/layers/mfdgp_hidden_layer.py:26-286,520-559``  (layer kernels, q(u) init, sample propagation,
  inducing inputs ``[Z, mean_{l-1}(Z)]``)
* ``mobocmf/models/mfdgp.py:22-262,290-317``               (model init, forward chain, predict,
  predict_for_acquisition)
* ``mobocmf/mlls/variational_elbo_mf.py:24-51``            (ELBO)
* ``mobocmf/acquisition_functions/JESMOC_MFDGP.py:38-52,118-135``  (JES term, coupled acquisition)
* ``mobocmf/util/blackbox_mfdgp_fitter.py:156-173,227-243,272-346`` (ELBO step, theta/omega factors,
  conditioned step)
* ``mobocmf/util/util.py:27-33``                           (triu_indices / compute_dist quirk Q2)
* upstream, from their public sources (gpytorch 1.11): ``UnwhitenedVariationalStrategy.forward``,
  ``CholeskyVariationalDistribution.forward``, ``kl_mvn_mvn``, ``Kernel.covar_dist``/``sq_dist``,
  ``RBFKernel``/``LinearKernel``/``ScaleKernel``/``ProductKernel``/``AdditiveKernel``,
  ``GaussianLikelihood.expected_log_prob``/``marginal``, ``MultivariateNormal.variance`` (min_variance clamp),
  ``Interval``/``Positive`` constraints, ``settings.variational_cholesky_jitter`` (1e-6 in fp64).

Parameters travel as a flat ``dict`` keyed with GPyTorch's ``state_dict`` names, values are
torch fp64 tensors; giving them ``requires_grad`` makes torch autograd the backward oracle.
"""
import math

import numpy as np
import torch

JITTER = 1e-6            # gpytorch.settings.variational_cholesky_jitter.value(torch.float64)   (quirk Q5)
MIN_VARIANCE = 1e-10     # gpytorch.settings.min_variance.value(torch.float64)                  (quirk Q9)
NOISE_LOWER = 1e-8       # models/mfdgp.py:116


# --------------------------------------------------------------------------------------------------
# constraints / transforms [upstream gpytorch.constraints, gpytorch.utils.transforms]
# --------------------------------------------------------------------------------------------------
def softplus(x):
    return torch.nn.functional.softplus(x)


def inv_softplus(x):
    return x + torch.log(-torch.expm1(-x))


def inv_sigmoid(x):
    return torch.log(x) - torch.log(1 - x)


def noise_value(raw_noise, lower, upper):
    """Interval(lower, upper).transform: lower + (upper - lower) * sigmoid(raw)  (models/mfdgp.py:116)."""
    return lower + (upper - lower) * torch.sigmoid(raw_noise)


# --------------------------------------------------------------------------------------------------
# kernels [upstream gpytorch.kernels] as instantiated at layers/mfdgp_hidden_layer.py:43-47,70-88,115
# --------------------------------------------------------------------------------------------------
def sq_dist(x1, x2):
    """Upstream ``gpytorch.kernels.kernel.sq_dist``: centred quadratic expansion + one matmul + clamp_min(0)."""
    adjustment = x1.mean(-2, keepdim=True)
    x1 = x1 - adjustment
    x2 = x2 - adjustment
    x1_norm = x1.pow(2).sum(dim=-1, keepdim=True)
    x1_pad = torch.ones_like(x1_norm)
    x2_norm = x2.pow(2).sum(dim=-1, keepdim=True)
    x2_pad = torch.ones_like(x2_norm)
    x1_ = torch.cat([-2.0 * x1, x1_norm, x1_pad], dim=-1)
    x2_ = torch.cat([x2, x2_pad, x2_norm], dim=-1)
    res = x1_.matmul(x2_.transpose(-2, -1))
    return res.clamp_min(0)


def rbf(x1, x2, lengthscale):
    """Upstream RBFKernel.forward (ARD branch): exp(-0.5 * sq_dist(x1/l, x2/l))."""
    return sq_dist(x1.div(lengthscale), x2.div(lengthscale)).div(-2).exp()


def _key_layer(l):
    return "hidden_layer_%d." % l


def layer_hypers(sd, l):
    """Constrained kernel hyper-parameters of layer l (softplus of the raw parameters)."""
    p = _key_layer(l) + "covar_module."
    if l == 0:
        return {"a": softplus(sd[p + "raw_outputscale"]),
                "ls": softplus(sd[p + "base_kernel.raw_lengthscale"])}
    return {"a1": softplus(sd[p + "kernels.0.kernels.0.raw_outputscale"]),
            "ls1": softplus(sd[p + "kernels.0.kernels.0.base_kernel.raw_lengthscale"]),
            "vlin": softplus(sd[p + "kernels.0.kernels.1.kernels.0.raw_variance"]).reshape(()),
            "af": softplus(sd[p + "kernels.0.kernels.1.kernels.1.raw_outputscale"]),
            "lsf": softplus(sd[p + "kernels.0.kernels.1.kernels.1.base_kernel.raw_lengthscale"]),
            "a2": softplus(sd[p + "kernels.1.raw_outputscale"]),
            "ls2": softplus(sd[p + "kernels.1.base_kernel.raw_lengthscale"])}


def layer_kernel(sd, l, A, B):
    """Dense cross-covariance K_l(A, B).  Layer 0: a*RBF; layer >= 1: k_x1*(k_lin + k_f) + k_x2
    (layers/mfdgp_hidden_layer.py:43-47,70-88,115).  The last column of A, B is the propagated f."""
    h = layer_hypers(sd, l)
    if l == 0:
        return h["a"] * rbf(A, B, h["ls"])
    xa, xb = A[:, :-1], B[:, :-1]
    fa, fb = A[:, -1:], B[:, -1:]
    k_x1 = h["a1"] * rbf(xa, xb, h["ls1"])
    k_lin = (fa * h["vlin"].sqrt()) @ (fb * h["vlin"].sqrt()).T       # upstream LinearKernel.forward
    k_f = h["af"] * rbf(fa, fb, h["lsf"])
    k_x2 = h["a2"] * rbf(xa, xb, h["ls2"])
    return k_x1 * (k_lin + k_f) + k_x2


def layer_kernel_diag(sd, l, A):
    """diag K_l(A, A) (upstream diag=True path: sq-dist diag is exactly zero, LinearKernel diag = v f^2)."""
    h = layer_hypers(sd, l)
    one = torch.ones(A.shape[0], dtype=A.dtype)
    if l == 0:
        return h["a"] * one
    f = A[:, -1]
    return h["a1"] * one * (h["vlin"] * f * f + h["af"] * one) + h["a2"] * one


# --------------------------------------------------------------------------------------------------
# q(u), unwhitened variational strategy [upstream], KL
# --------------------------------------------------------------------------------------------------
def psd_safe_cholesky(A, jitter=1e-8, max_tries=3):
    """Upstream linear_operator psd_safe_cholesky: retry with growing diagonal jitter."""
    L, info = torch.linalg.cholesky_ex(A)
    if not bool(info.any()):
        return L
    if torch.isnan(A).any():
        raise RuntimeError("NaNError: cholesky of a matrix with NaNs")
    Aprime = A.clone()
    jitter_prev = 0.0
    for i in range(max_tries):
        jitter_new = jitter * (10 ** i)
        Aprime.diagonal().add_(jitter_new - jitter_prev)
        jitter_prev = jitter_new
        L, info = torch.linalg.cholesky_ex(Aprime)
        if not bool(info.any()):
            return L
    raise RuntimeError("NotPSDError: matrix not positive definite after jitter")


def variational_q(sd, l):
    """CholeskyVariationalDistribution.forward: (m, L_q = tril(chol_variational_covar))."""
    p = _key_layer(l) + "variational_strategy._variational_distribution."
    return sd[p + "variational_mean"], torch.tril(sd[p + "chol_variational_covar"])


def layer_inducing_points(sd, l):
    """MFDGUnwhitenedVariationalStrategy.inducing_points (layers/mfdgp_hidden_layer.py:542-559):
    Z_l = [Z[:, :d], previous_layer(Z[:, :d]).mean], recomputed on every access, gradient flows.

    Reference depth is L = 2 (fact F3).  L >= 3 extension (parity unpinned): a layer l-1 >= 1 evaluated at its
    own inducing inputs returns m_{l-1} (the quirk-Q4 shortcut), so Z_l = [Z, m_{l-1}] when Z is shared."""
    Zo = sd[_key_layer(l) + "variational_strategy.inducing_points"]
    if l == 0:
        return Zo
    zx = Zo[:, :-1]
    if l - 1 >= 1:
        zx_prev = sd[_key_layer(l - 1) + "variational_strategy.inducing_points"][:, :-1]
        if not torch.equal(zx, zx_prev):
            raise NotImplementedError("L>=3 extension requires shared inducing inputs")
        mean_prev = variational_q(sd, l - 1)[0]
    else:
        mean_prev, _ = layer_q(sd, 0, zx, training=True)
    return torch.cat((zx, mean_prev[:, None]), 1)


def prior_cholesky(sd, l, Z=None, jitter=JITTER):
    if Z is None:
        Z = layer_inducing_points(sd, l)
    M = Z.shape[0]
    P = layer_kernel(sd, l, Z, Z) + jitter * torch.eye(M, dtype=Z.dtype)
    return psd_safe_cholesky(P)


def layer_q(sd, l, X, training=True, jitter=JITTER, literal_eval_cov=False):
    """q(f_l(X)) = (mean, raw variance) — upstream UnwhitenedVariationalStrategy.forward.

    X already carries the propagated f column for l >= 1.  ``training`` selects the clamp(k_xx - q, 0) branch
    (module.training) versus the eval branch (no inner clamp, solve-based q).  The returned variance is RAW;
    callers apply the ``.variance`` floor (quirk Q9) with :func:`read_variance`."""
    Z = layer_inducing_points(sd, l)
    m, Lq = variational_q(sd, l)
    if torch.equal(X, Z):                                   # quirk Q4 shortcut: N(m, L_q L_q^T)
        return m, (Lq ** 2).sum(-1)
    Lp = prior_cholesky(sd, l, Z, jitter)
    Kzx = layer_kernel(sd, l, Z, X)
    kxx = layer_kernel_diag(sd, l, X)
    left = torch.cat([m[:, None], Lq], -1)                  # [mean_diff, root_variational_covar]
    A = torch.cholesky_solve(Kzx, Lp)                       # P^-1 K_zx
    inv_products = left.mT @ A                              # (1+M, R)
    mean = inv_products[0]
    if training:
        interp = (torch.linalg.solve_triangular(Lp, Kzx, upper=False) ** 2).sum(0)
        data_var = (kxx - interp).clamp(0, math.inf)
    elif literal_eval_cov:                                  # the literal R x R eval branch (small R only)
        Kxx = layer_kernel(sd, l, X, X)
        data_var = (Kxx + Kzx.mT.mul(-1) @ A).diagonal()
    else:
        data_var = kxx - (Kzx * A).sum(0)
    var = (inv_products[1:] ** 2).sum(0) + data_var
    return mean, var


def read_variance(v):
    """MultivariateNormal.variance: clamp_min(settings.min_variance) on every read (quirk Q9)."""
    return v.clamp_min(MIN_VARIANCE)


def kl_layer(sd, l, jitter=JITTER):
    """KL(q(u_l) || p(u_l)) — upstream kl_mvn_mvn through the Cholesky path."""
    Lp = prior_cholesky(sd, l, None, jitter)
    m, Lq = variational_q(sd, l)
    M = m.shape[0]
    rhs = torch.cat([m[:, None], Lq], -1)
    R = torch.linalg.solve_triangular(Lp, rhs, upper=False)
    trace_plus_inv_quad = (R ** 2).sum()
    logdet_prior = Lp.diagonal().pow(2).log().sum()
    logdet_q = Lq.diagonal().pow(2).log().sum()
    return 0.5 * (logdet_prior - logdet_q + trace_plus_inv_quad - float(M))


def kl_divergence(sd, num_layers, jitter=JITTER):
    """DeepGP variational_strategy.kl_divergence(): sum over layers, each counted once (quirk Q12)."""
    return sum(kl_layer(sd, l, jitter) for l in range(num_layers))


# --------------------------------------------------------------------------------------------------
# model chain: MFDGP.forward / predict / predict_for_acquisition  (models/mfdgp.py:174-262)
# --------------------------------------------------------------------------------------------------
def mfdgp_forward(sd, num_layers, x, eps=None, samples=None, training=True, eval_mode=False,
                  max_fidelity=None, only_hf=False, jitter=JITTER):
    """MFDGP.forward (models/mfdgp.py:174-196) with the layer __call__ of layers/...py:245-286.

    eps[l]      : train-mode normals of layer l >= 1 (reference: float32 ``torch.normal`` of shape (1,B), Q6).
    samples[l]  : eval-mode fixed normals (S,1) owned by layer l >= 1 (Q7); row i*S+s uses samples[l][s].
    Returns a list of (mean, raw_variance); layer 0 has shape (1,R), layers >= 1 shape (R,) (quirk Q3)."""
    nl = num_layers if max_fidelity is None else max_fidelity + 1
    outs = []
    for l in range(nl):
        if l == 0:
            mean, var = layer_q(sd, 0, x, training, jitter)
            outs.append((mean[None, :], var[None, :]))
            continue
        pm, pv = outs[-1]
        if only_hf:                                         # models/mfdgp.py:189-190, layers/...py:280
            f = (pm * 0.0).reshape(-1, 1)
        elif eval_mode:                                     # layers/...py:263-270
            S = samples[l].shape[0]
            n = pm.numel() // S
            rep = samples[l].to(x.dtype).reshape(1, S, 1).repeat_interleave(n, 0).reshape(pm.shape)
            f = (pm + torch.sqrt(read_variance(pv)) * rep).reshape(x.shape[0], 1)
        else:                                               # layers/...py:274
            e = eps[l].to(x.dtype).reshape(pm.shape)
            f = (e * torch.sqrt(read_variance(pv)) + pm).reshape(-1, 1)
        xa = torch.cat([x, f], dim=-1)
        outs.append(layer_q(sd, l, xa, training, jitter))
    return outs


def likelihood_noise(sd, l, noise_upper, noise_lower=NOISE_LOWER):
    raw = sd["hidden_layer_likelihood_%d.noise_covar.raw_noise" % l]
    return noise_value(raw, noise_lower[l] if isinstance(noise_lower, (list, tuple)) else noise_lower,
                       noise_upper[l]).reshape(())


def predict(sd, num_layers, noise_upper, test_x, fidelity_layer=0, noise_lower=NOISE_LOWER, **kw):
    """MFDGP.predict (models/mfdgp.py:220-235): likelihood_f(q(f_f)) -> (mean, variance + noise)."""
    assert 0 <= fidelity_layer < num_layers
    outs = mfdgp_forward(sd, num_layers, test_x, max_fidelity=fidelity_layer, **kw)
    mean, var = outs[fidelity_layer]
    noise = likelihood_noise(sd, fidelity_layer, noise_upper, noise_lower)
    return mean, read_variance(var + noise)


def predict_for_acquisition(sd, num_layers, noise_upper, samples, test_x, fidelity_layer=0, training=False,
                            only_hf=False, noise_lower=NOISE_LOWER, jitter=JITTER):
    """MFDGP.predict_for_acquisition (models/mfdgp.py:237-262): tile x S, eval-mode pass, moment match."""
    if test_x.dim() > 2:
        assert test_x.shape[1] == 1
        test_x = test_x[:, 0, :]
    S = samples[1].shape[0] if num_layers > 1 else 1
    x_tile = test_x.repeat_interleave(S, 0)
    outs = mfdgp_forward(sd, num_layers, x_tile, samples=samples, training=training, eval_mode=True,
                         max_fidelity=fidelity_layer, only_hf=only_hf, jitter=jitter)
    mus_tilde, var_raw = outs[fidelity_layer]
    noise = likelihood_noise(sd, fidelity_layer, noise_upper, noise_lower)
    vars_tilde = read_variance(var_raw + noise)
    n = test_x.shape[0]
    mus = torch.mean(torch.reshape(mus_tilde, (n, S)), 1)
    second_moment = torch.mean(torch.reshape(vars_tilde + mus_tilde ** 2, (n, S)), 1)
    return mus, second_moment - mus ** 2


# --------------------------------------------------------------------------------------------------
# ELBO  (mlls/variational_elbo_mf.py:24-51) and the ELBO step (util/blackbox_mfdgp_fitter.py:156-173)
# --------------------------------------------------------------------------------------------------
def expected_log_prob(target, mean, var, noise):
    """Upstream GaussianLikelihood.expected_log_prob."""
    res = ((target - mean).square() + var) / noise + noise.log() + math.log(2 * math.pi)
    return res.mul(-0.5)


def elbo(sd, num_layers, noise_upper, outs, target, fidelities, num_data, include_kl_term=True,
         noise_lower=NOISE_LOWER, jitter=JITTER):
    """VariationalELBOMF.forward: (data - KL*B/N, KL*B/N) or the data term alone."""
    assert target.shape[0] <= target.shape[1]
    num_batch = target.shape[1]
    data_term = 0.0
    for i in range(num_layers):
        if (fidelities == i).sum() != 0:
            mean, var = outs[i]
            noise = likelihood_noise(sd, i, noise_upper, noise_lower)
            ell = expected_log_prob(target, mean, read_variance(var), noise)
            data_term = data_term + ell[fidelities.T == i].sum()
    if not include_kl_term:
        return data_term
    kl = kl_divergence(sd, num_layers, jitter)
    return data_term - kl * num_batch / num_data, kl * num_batch / num_data


def elbo_step_loss(sd, num_layers, noise_upper, x_batch, y_batch, fid_batch, eps, num_data, **kw):
    """Body of ``_update_model`` up to the loss: -ELBO and the scaled KL (fitter.py:163-167)."""
    outs = mfdgp_forward(sd, num_layers, x_batch, eps=eps, training=True,
                         jitter=kw.get("jitter", JITTER), only_hf=kw.get("only_hf", False))
    e, kl = elbo(sd, num_layers, noise_upper, outs, y_batch.T, fid_batch, num_data,
                 noise_lower=kw.get("noise_lower", NOISE_LOWER), jitter=kw.get("jitter", JITTER))
    return -e, kl


def elbo_step_loss_multisample(sd, num_layers, noise_upper, x_batch, y_batch, fid_batch, eps_list, num_data,
                               **kw):
    """S>1 training (fact F4: not in the reference; defined as the mean of S single-sample ELBOs with the
    given eps).  eps_list[s][l] are the normals of sample s, layer l."""
    losses, kls = zip(*[elbo_step_loss(sd, num_layers, noise_upper, x_batch, y_batch, fid_batch, e, num_data,
                                       **kw) for e in eps_list])
    return sum(losses) / len(losses), kls[0]


def elbo_step_loss_tiled(sd, num_layers, noise_upper, x_batch, y_batch, fid_batch, eps, num_data, S,
                         noise_lower=NOISE_LOWER, jitter=JITTER):
    """The same S-sample ELBO evaluated the way a CPU user would: layer 0 once on the B rows, upper layers on the
    B*S tiled rows (row b*S+s), data terms averaged over s.  eps[l]: (B*S,) normals of layer l.  Used as the CPU
    baseline of config C4 and cross-checked against :func:`elbo_step_loss_multisample`."""
    B = x_batch.shape[0]
    mean, var = layer_q(sd, 0, x_batch, True, jitter)
    outs = [(mean, var)]
    x_tile = x_batch.repeat_interleave(S, 0)
    pm, pv = mean.repeat_interleave(S, 0), var.repeat_interleave(S, 0)
    for l in range(1, num_layers):
        f = (eps[l].to(x_batch.dtype).reshape(-1) * torch.sqrt(read_variance(pv)) + pm).reshape(-1, 1)
        pm, pv = layer_q(sd, l, torch.cat([x_tile, f], -1), True, jitter)
        outs.append((pm, pv))
    data_term = 0.0
    y = y_batch.reshape(-1)
    for l in range(num_layers):
        mask = fid_batch.reshape(-1) == l
        if mask.sum() != 0:
            noise = likelihood_noise(sd, l, noise_upper, noise_lower)
            m_, v_ = outs[l]
            if l == 0:
                ell = expected_log_prob(y, m_, read_variance(v_), noise)
            else:
                ell = expected_log_prob(y.repeat_interleave(S, 0), m_, read_variance(v_), noise).reshape(B, S).mean(1)
            data_term = data_term + ell[mask].sum()
    kl = kl_divergence(sd, num_layers, jitter)
    return -(data_term - kl * B / num_data), kl * B / num_data


def adam_step(params, grads, state, lr, betas=(0.9, 0.999), eps=1e-8):
    """torch.optim.Adam defaults (fitter.py:126,132,259), restated for the fused-Adam parity test."""
    state["step"] = state.get("step", 0) + 1
    t = state["step"]
    out = []
    for i, (p, g) in enumerate(zip(params, grads)):
        m = state.setdefault(("m", i), torch.zeros_like(p))
        v = state.setdefault(("v", i), torch.zeros_like(p))
        m.mul_(betas[0]).add_(g, alpha=1 - betas[0])
        v.mul_(betas[1]).addcmul_(g, g, value=1 - betas[1])
        bc1 = 1 - betas[0] ** t
        bc2 = 1 - betas[1] ** t
        denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
        out.append(p - (lr / bc1) * m / denom)
    return out


# --------------------------------------------------------------------------------------------------
# JES acquisition  (acquisition_functions/JESMOC_MFDGP.py:38-52,118-135)
# --------------------------------------------------------------------------------------------------
def jes(var_uncond, var_cond):
    """_JES_MFDGP.forward: 0.5 * clamp(log v_u - log v_c, min=0)."""
    return 0.5 * torch.clamp(torch.log(var_uncond) - torch.log(var_cond), min=0.0)


def jes_mfdgp(model_u, model_c, X, fidelity):
    """One black box: model_* = dict(sd=, num_layers=, noise_upper=, samples=, [only_hf, noise_lower])."""
    def pfa(mod):
        return predict_for_acquisition(mod["sd"], mod["num_layers"], mod["noise_upper"], mod["samples"], X,
                                       fidelity, training=False, only_hf=mod.get("only_hf", False),
                                       noise_lower=mod.get("noise_lower", NOISE_LOWER))
    _, vu = pfa(model_u)
    _, vc = pfa(model_c)
    return jes(vu, vc)


def coupled_acq(models_u, models_c, X, fidelity, float32_accumulator=False):
    """JESMOC_MFDGP.coupled_acq: sum over objectives + constraints.  The reference accumulates fp64 terms
    into a float32 tensor in place (quirk Q8); parity is asserted on the fp64 sum unless asked otherwise."""
    acq = torch.zeros(X.shape[0], dtype=torch.float32 if float32_accumulator else torch.float64)
    for mu, mc in zip(models_u, models_c):
        acq += jes_mfdgp(mu, mc, X.double(), fidelity)
    return acq


# --------------------------------------------------------------------------------------------------
# conditioned training  (util/blackbox_mfdgp_fitter.py:227-243,272-346)
# --------------------------------------------------------------------------------------------------
_std_normal = torch.distributions.normal.Normal(0.0, 1.0)


def loss_theta_factors(cs_mean, cs_var, threshold, eps=1e-8):
    gamma = (cs_mean - threshold) / torch.sqrt(cs_var)
    cdf = _std_normal.cdf(gamma)
    return torch.sum(np.log(1.0 - eps) * cdf + np.log(eps) * (1.0 - cdf))


def loss_omega_factors(fs_mean, fs_var, cs_mean, cs_var, pareto_front, thresholds_cons, eps=1e-8):
    gamma_c = (cs_mean - thresholds_cons[:, None]) / torch.sqrt(cs_var)
    gamma_f_star = (pareto_front[:, :, None] - fs_mean) / torch.sqrt(fs_var)
    prod = torch.prod(_std_normal.cdf(gamma_c), 0) * torch.prod(_std_normal.cdf(gamma_f_star), 1)
    return torch.sum(np.log(eps) * prod + np.log(1 - eps) * (1.0 - prod))


def conditioned_step_loss(objs, cons, pareto_set, pareto_front, thresholds_cons, x_tilde, eps_factor=1e-8):
    """Body of ``_update_conditioned_models`` (fitter.py:272-346) up to the loss.

    objs / cons: lists of dict(sd, num_layers, noise_upper, num_data, batch=(x,y,fid), eps_batch, eps_pareto,
    eps_tilde).  Each MFDGP forward is a single-sample stochastic pass with the given normals."""
    loss = 0.0

    def fwd(mod, x, eps):
        return mfdgp_forward(mod["sd"], mod["num_layers"], x, eps=eps, training=True)

    for i, mod in enumerate(objs):
        xb, yb, fb = mod["batch"]
        outs = fwd(mod, xb, mod["eps_batch"])
        e = elbo(mod["sd"], mod["num_layers"], mod["noise_upper"], outs, yb.T, fb, mod["num_data"])[0]
        loss = loss + -e / xb.shape[0] * mod["num_data"]
        outs = fwd(mod, pareto_set, mod["eps_pareto"])
        pf = torch.ones(pareto_front.shape[0], 1) * (mod["num_layers"] - 1)
        loss = loss + -elbo(mod["sd"], mod["num_layers"], mod["noise_upper"], outs,
                            pareto_front[:, i:i + 1].T, pf, mod["num_data"], include_kl_term=False)
    for k, mod in enumerate(cons):
        xb, yb, fb = mod["batch"]
        outs = fwd(mod, xb, mod["eps_batch"])
        e = elbo(mod["sd"], mod["num_layers"], mod["noise_upper"], outs, yb.T, fb, mod["num_data"])[0]
        loss = loss + -e / xb.shape[0] * mod["num_data"]
        mean, var = fwd(mod, pareto_set, mod["eps_pareto"])[mod["num_layers"] - 1]
        loss = loss + -loss_theta_factors(mean, read_variance(var), thresholds_cons[k], eps_factor)
    fm, fv, cm, cv = [], [], [], []
    for mod in objs:
        mean, var = fwd(mod, x_tilde, mod["eps_tilde"])[mod["num_layers"] - 1]
        fm.append(mean.reshape(1, -1)); fv.append(read_variance(var).reshape(1, -1))
    for mod in cons:
        mean, var = fwd(mod, x_tilde, mod["eps_tilde"])[mod["num_layers"] - 1]
        cm.append(mean.reshape(1, -1)); cv.append(read_variance(var).reshape(1, -1))
    z = torch.zeros(0, x_tilde.shape[0], dtype=torch.double)
    loss = loss + -loss_omega_factors(torch.cat(fm, 0) if fm else z, torch.cat(fv, 0) if fv else z,
                                      torch.cat(cm, 0) if cm else z, torch.cat(cv, 0) if cv else z,
                                      pareto_front, thresholds_cons, eps_factor)
    return loss


# --------------------------------------------------------------------------------------------------
# model construction  (models/mfdgp.py:22-151,290-317; layers/mfdgp_hidden_layer.py:26-161; util/util.py:27-33)
# --------------------------------------------------------------------------------------------------
def triu_indices(n, offset=0):
    rows, cols = torch.triu_indices(n, n, offset=offset)
    return torch.stack((rows, cols), dim=0)


def compute_dist(x):
    return torch.sum(x ** 2, 1, keepdims=True) - 2.0 * x.mm(x.T) + torch.sum(x ** 2, 1, keepdims=True).T


def init_lengthscale_median(inputs):
    """get_init_lengthscale(TL.MEDIAN) including quirk Q2 (row indexing with a (2,K) LongTensor)."""
    d = compute_dist(inputs)
    return torch.sqrt(torch.median(d[triu_indices(inputs.shape[0], 1)]))


def initial_inducing_points_and_values(x_train, y_train, fidelities, layer, only_hf=False):
    """find_good_initial_inducing_points_and_values (models/mfdgp.py:290-317); float32 values (quirk Q1)."""
    sel = fidelities[:, 0] == layer
    inducing_points = x_train[sel, :] if only_hf else x_train
    inducing_values = torch.zeros(inducing_points.shape[0])
    xs, ys = x_train[sel, :], y_train[sel, :]
    for i in range(inducing_points.shape[0]):
        tmp = torch.cat((xs, inducing_points[i:i + 1, :]), 0)
        to_sel = torch.argmin(compute_dist(tmp)[0:tmp.shape[0] - 1, tmp.shape[0] - 1])
        inducing_values[i] = ys[to_sel]
    if layer != 0:
        inducing_points = torch.cat((inducing_points, inducing_values[:, None]), 1)
    return inducing_points, inducing_values


def init_state_dict(x_train, y_train, fidelities, num_fidelities, type_lengthscale="median",
                    num_samples_for_acquisition=25, only_hf=False, generator=None):
    """MFDGP.__init__ + MFDGPHiddenLayer.__init__ + ``.double()`` (fitter.py:32): parameters are created in
    float32 and then widened, so every initial value is a float32 rounding (quirk Q1 generalised).

    Returns (sd, noise_lower[L], noise_upper[L], samples[L])."""
    d = x_train.shape[-1]
    y_high_std = np.std(y_train[(fidelities == num_fidelities - 1).flatten()].numpy())
    sd, samples, lowers, uppers = {}, [], [], []
    f32 = torch.float32

    def raw_pos(value, shape):
        v = torch.as_tensor(value).to(f32)
        return inv_softplus(v).expand(shape).clone().double()

    for l in range(num_fidelities):
        Z, vals = initial_inducing_points_and_values(x_train, y_train, fidelities, l, only_hf)
        xs = x_train[(fidelities == l).flatten(), :]
        if type_lengthscale == "median":
            ls0 = init_lengthscale_median(xs)
        elif type_lengthscale == "ones":
            ls0 = torch.ones(d)
        else:
            ls0 = torch.as_tensor(0.01 * np.ones(d))
        p = _key_layer(l)
        c = p + "covar_module."
        if l == 0:
            sd[c + "raw_outputscale"] = raw_pos(1.0, ())
            sd[c + "base_kernel.raw_lengthscale"] = raw_pos(ls0, (1, d))
        else:
            a1, af, a2, vl = (0.0, 0.0, 1.0, 0.0) if only_hf else (1.0, 1.0, 0.01, 1.0)
            sd[c + "kernels.0.kernels.0.raw_outputscale"] = raw_pos(a1, ())
            sd[c + "kernels.0.kernels.0.base_kernel.raw_lengthscale"] = raw_pos(ls0 * 10.0, (1, d))
            sd[c + "kernels.0.kernels.1.kernels.0.raw_variance"] = raw_pos(vl, (1, 1))
            sd[c + "kernels.0.kernels.1.kernels.1.raw_outputscale"] = raw_pos(af, ())
            sd[c + "kernels.0.kernels.1.kernels.1.base_kernel.raw_lengthscale"] = raw_pos(1.0, (1, 1))
            sd[c + "kernels.1.raw_outputscale"] = raw_pos(a2, ())
            sd[c + "kernels.1.base_kernel.raw_lengthscale"] = raw_pos(ls0, (1, d))
        M = Z.shape[0]
        sd[p + "variational_strategy.inducing_points"] = Z.double()
        sd[p + "variational_strategy.variational_params_initialized"] = torch.tensor(1)
        q = p + "variational_strategy._variational_distribution."
        sd[q + "variational_mean"] = vals.to(f32).double()
        if l == num_fidelities - 1:                          # layers/...py:131-132
            cov = layer_kernel(sd, l, Z.double(), Z.double()) * (1e-2 * y_high_std ** 2) ** 2
            sd[q + "chol_variational_covar"] = psd_safe_cholesky(cov).to(f32).double()
        else:                                                # layers/...py:134
            sd[q + "chol_variational_covar"] = psd_safe_cholesky(torch.eye(M) * 1e-8).to(f32).double()
        y_std = np.std(y_train[(fidelities == l).flatten()].numpy())
        lower = torch.as_tensor(NOISE_LOWER).to(f32)
        upper = torch.as_tensor(0.1 * y_std).to(f32)
        init_noise = torch.as_tensor(1e-2 * y_high_std if l == num_fidelities - 1 else 1e-6).to(f32)
        raw = inv_sigmoid((init_noise - lower) / (upper - lower))
        sd["hidden_layer_likelihood_%d.noise_covar.raw_noise" % l] = raw.reshape(1).double()
        lowers.append(float(lower.double()))
        uppers.append(float(upper.double()))
        s = torch.normal(mean=torch.zeros([num_samples_for_acquisition]),
                         std=torch.ones([num_samples_for_acquisition]), generator=generator)[:, None]
        samples.append(s)
    return sd, lowers, uppers, samples
