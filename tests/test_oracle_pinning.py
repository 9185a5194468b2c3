"""Pins the oracle to closed forms computed INDEPENDENTLY of its own code path (the reference cannot run here and ships
no golden vectors, SURVEY.md section 8c, so the oracle is 'parity unpinned' with respect to the reference itself):

* KL(q(u) || p(u)) against torch.distributions' MultivariateNormal KL;
* the sparse-GP predictive against the dense joint-Gaussian conditional in numpy longdouble:
      mean = k^T P^-1 m,   var = k_xx - k^T P^-1 k + k^T P^-1 S P^-1 k     (S = L_q L_q^T, P = K_zz + jitter I)
  with explicit extended-precision inverses instead of Cholesky solves, and direct-difference distances instead of the
  quadratic expansion;
* the expected log-likelihood against numerical Gauss-Hermite quadrature of E_q[log N(y | f, s2)];
* the moment matching of predict_for_acquisition against the mixture-of-Gaussians moments.
"""
import math

import numpy as np
import torch

from oracle import mfdgp_oracle as O
from tests.helpers import random_state

LD = np.longdouble


def dense_kernel_longdouble(sd, l, A, B):
    """Direct-difference evaluation of the layer covariance in extended precision."""
    h = {k: v.detach().numpy().astype(LD) for k, v in O.layer_hypers(sd, l).items()}
    A, B = A.numpy().astype(LD), B.numpy().astype(LD)

    def rbf(xa, xb, ls):
        d = (xa[:, None, :] - xb[None, :, :]) / ls.reshape(1, 1, -1)
        return np.exp(-0.5 * (d ** 2).sum(-1))
    if l == 0:
        return h["a"] * rbf(A, B, h["ls"])
    xa, xb, fa, fb = A[:, :-1], B[:, :-1], A[:, -1:], B[:, -1:]
    return h["a1"] * rbf(xa, xb, h["ls1"]) * (h["vlin"] * fa @ fb.T + h["af"] * rbf(fa, fb, h["lsf"])) + \
        h["a2"] * rbf(xa, xb, h["ls2"])


def test_kl_matches_torch_distributions():
    for M, d, L, ls in [(10, 2, 2, 0.4), (17, 3, 3, 0.3)]:
        sd, _ = random_state(M, d, L, seed=M, ls=ls)
        for l in range(L):
            Z = O.layer_inducing_points(sd, l)
            P = O.layer_kernel(sd, l, Z, Z) + O.JITTER * torch.eye(M, dtype=torch.float64)
            m, Lq = O.variational_q(sd, l)
            q = torch.distributions.MultivariateNormal(m, scale_tril=Lq)
            p = torch.distributions.MultivariateNormal(torch.zeros(M, dtype=torch.float64), covariance_matrix=P)
            ref = torch.distributions.kl_divergence(q, p)
            got = O.kl_layer(sd, l)
            assert abs(float(got - ref)) < 1e-9 * abs(float(ref)), (l, float(got), float(ref))


def test_predictive_matches_dense_conditional_in_extended_precision():
    M, d, L = 14, 2, 2
    sd, _ = random_state(M, d, L, seed=5, ls=0.45)
    g = torch.Generator().manual_seed(1)
    x = torch.rand(9, d, generator=g, dtype=torch.float64)
    f = torch.randn(9, 1, generator=g, dtype=torch.float64)
    for l in range(L):
        X = x if l == 0 else torch.cat([x, f], 1)
        Z = O.layer_inducing_points(sd, l)
        m, Lq = O.variational_q(sd, l)
        P = dense_kernel_longdouble(sd, l, Z, Z) + LD(O.JITTER) * np.eye(M, dtype=LD)
        Kzx = dense_kernel_longdouble(sd, l, Z, X)
        kxx = np.diag(dense_kernel_longdouble(sd, l, X, X))
        # extended-precision inverse by Gauss-Jordan (numpy.linalg has no longdouble path)
        A = np.concatenate([P, np.eye(M, dtype=LD)], 1)
        for i in range(M):
            piv = i + int(np.argmax(np.abs(A[i:, i])))
            A[[i, piv]] = A[[piv, i]]
            A[i] = A[i] / A[i, i]
            for r in range(M):
                if r != i:
                    A[r] = A[r] - A[r, i] * A[i]
        Pinv = A[:, M:]
        S = Lq.numpy().astype(LD) @ Lq.numpy().astype(LD).T
        a = Pinv @ Kzx
        mean = a.T @ m.numpy().astype(LD)
        var = kxx - (Kzx * a).sum(0) + ((S @ a) * a).sum(0)
        mo, vo = O.layer_q(sd, l, X, training=True)
        cond = float(np.linalg.cond(P.astype(np.float64)))
        tol = max(1e-10, 50 * 2.2e-16 * cond)
        assert np.abs(mo.numpy() - mean.astype(np.float64)).max() < tol * max(1.0, np.abs(mean).max()), (l, cond)
        assert np.abs(vo.numpy() - var.astype(np.float64)).max() < tol * max(1.0, np.abs(var).max()), (l, cond)
        me, ve = O.layer_q(sd, l, X, training=False)      # eval branch: same numbers when nothing is clamped
        assert torch.allclose(me, mo, rtol=1e-9, atol=1e-12) and torch.allclose(ve, vo, rtol=1e-7, atol=1e-10)


def test_expected_log_prob_matches_quadrature():
    y, mu, var, s2 = 0.3, -0.2, 0.7, 0.05
    t, w = np.polynomial.hermite_e.hermegauss(60)
    f = mu + math.sqrt(var) * t
    quad = (w * (-0.5 * ((y - f) ** 2 / s2 + math.log(s2) + math.log(2 * math.pi)))).sum() / math.sqrt(2 * math.pi)
    got = O.expected_log_prob(*[torch.tensor(v, dtype=torch.float64) for v in (y, mu, var, s2)])
    assert abs(float(got) - quad) < 1e-12 * abs(quad)


def test_moment_matching_is_the_mixture_moments():
    M, d, L = 10, 2, 2
    sd, noise_upper = random_state(M, d, L, seed=2, ls=0.5)
    g = torch.Generator().manual_seed(3)
    samples = [torch.randn(6, 1, generator=g) for _ in range(L)]
    X = torch.rand(4, d, generator=g, dtype=torch.float64)
    mu, var = O.predict_for_acquisition(sd, L, noise_upper, samples, X, 1)
    # by hand: S single-sample passes through the chain with f = mu0 + sqrt(v0) * sample_s
    m0, v0 = O.layer_q(sd, 0, X, training=False)
    mus, vs = [], []
    for s in range(6):
        f = m0 + torch.sqrt(O.read_variance(v0)) * samples[1][s].double()
        m1, v1 = O.layer_q(sd, 1, torch.cat([X, f[:, None]], 1), training=False)
        mus.append(m1)
        vs.append(O.read_variance(v1 + O.likelihood_noise(sd, 1, noise_upper)))
    mus, vs = torch.stack(mus), torch.stack(vs)
    mix_mean = mus.mean(0)
    mix_var = (vs + mus ** 2).mean(0) - mix_mean ** 2
    assert torch.allclose(mu, mix_mean, rtol=1e-12, atol=1e-14) and torch.allclose(var, mix_var, rtol=1e-10, atol=1e-14)
