"""Random-Fourier-feature function samples of the MFDGP layers — the B200 side of
``MFDGPHiddenLayer._sample_from_posterior(_layer0)`` / ``_sample_from_prior(_layer0)``
(``mobocmf/layers/mfdgp_hidden_layer.py:288-514``) and of the closures they return.

* The random draws (W, b, the standard normals behind theta) come from numpy's global generator in the reference's
  order, so ``np.random.seed(s)`` gives the same spectral frequencies and phases as the reference.
* The posterior weights (``_rff_sample_posterior_weights``, :298-309: two Cholesky factorisations and an inverse of a
  (k F) x (k F) matrix, F = 500, k = 1 or 3; once per layer and Pareto sample) are computed on the layer's GPU with
  torch.linalg in fp64 — a one-off library call, not the data-parallel part.
* A sample is an ``RFFSample``: called like the reference's ``wrapper(x, gradient=False)`` it evaluates the whole
  fidelity chain below it at all rows of x in ONE launch of ``mobo_rff_eval`` (csrc/rff.cu) — the data-parallel part
  (MOOP evaluates every sample on a 1000 d^2-point grid).  No CPU fallback: CPU inputs are copied to the GPU and back.
"""
import ctypes
import math

import numpy as np
import torch

from . import _lib


def posterior_weights(y, S, Phi, randomness, sigma2=1e-6):
    """theta ~ N(A^-1 Phi y, sigma2 A^-1 + A^-1 Phi S Phi^T A^-1), A = Phi Phi^T + sigma2 I  (:298-309); device fp64."""
    nfeat = Phi.shape[0]
    eye = torch.eye(nfeat, dtype=Phi.dtype, device=Phi.device)
    A = Phi @ Phi.T + sigma2 * eye
    cA = torch.linalg.cholesky(A)                               # lower; the reference's upper factor transposed
    A_inv = torch.cholesky_solve(eye, cA)
    mean = torch.cholesky_solve((Phi @ y).unsqueeze(1), cA).squeeze(1)
    AiP = A_inv @ Phi
    cov = sigma2 * A_inv + AiP @ S @ AiP.T
    cov = 0.5 * (cov + cov.T)
    U = torch.linalg.cholesky(cov, upper=True)
    return mean + (randomness @ U)


def _phi(x, W, b, alpha, nF):
    return math.sqrt(2.0 * alpha / nF) * torch.cos(W @ x.T + b)


class RFFSample(object):
    """Callable function sample of layer ``len(chain) - 1`` (and, through ``all_layers``, of every layer below it).

    chain[l]: dict with the raw draws as device tensors (W / b / theta for layer 0; W_x1, W_f, W_x2, b_x1, b_x2, theta
    for layer >= 1) and the amplitudes; packed once into the parameter blocks of include/mobocmf_b200.h."""

    def __init__(self, chain, input_dim, nFeatures, device):
        self.chain = chain
        self.d = int(input_dim)
        self.nF = int(nFeatures)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("mobocmf_b200: RFF function samples live on a CUDA device (no CPU fallback)")
        self._params, scales = [], []
        for s in chain:
            F = self.nF
            if s["kind"] == 0:
                blk = torch.cat([s["W"].reshape(-1), s["b"].reshape(-1), s["theta"].reshape(-1)])
                scales += [math.sqrt(2.0 * s["alpha"] / F), 0.0, 0.0]
            else:
                blk = torch.cat([s["W_x1"].reshape(-1), s["W_f"].reshape(-1), s["W_x2"].reshape(-1),
                                 s["b_x1"].reshape(-1), s["b_x2"].reshape(-1), s["theta"].reshape(-1)])
                scales += [math.sqrt(2.0 * s["alpha_x1"] / F) * math.sqrt(s["nu_lin"]),
                           math.sqrt(2.0 * s["alpha_x1f"] / F), math.sqrt(2.0 * s["alpha_x2"] / F)]
            self._params.append(blk.to(device=self.device, dtype=torch.float64).contiguous())
        self._scale_values = scales
        self._scales = self._ptrs = None      # ctypes views, built on first use (not picklable / deep-copyable)

    def __getstate__(self):
        # the fitter that stores the samples is deep-copied (copy_uncond, util/blackbox_mfdgp_fitter.py:383)
        state = self.__dict__.copy()
        state["_scales"] = state["_ptrs"] = None
        return state

    def _ctypes_views(self):
        if self._ptrs is None:
            self._scales = (ctypes.c_double * len(self._scale_values))(*self._scale_values)
            self._ptrs = _lib.ptr_array(self._params)
        return self._ptrs, self._scales

    @property
    def num_layers(self):
        return len(self.chain)

    def _run(self, x, want_grad):
        x = x.to(device=self.device, dtype=torch.float64)
        if x.ndim == 1:
            x = x[None, :]
        x = x.contiguous()
        assert x.shape[1] == self.d, (x.shape, self.d)
        n = x.shape[0]
        f = torch.empty(self.num_layers, n, dtype=torch.float64, device=self.device)
        g = torch.empty(n, self.d, dtype=torch.float64, device=self.device) if want_grad else None
        ptrs, scales = self._ctypes_views()
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().mobo_rff_eval(self.num_layers, self.d, self.nF, ptrs, scales,
                                                 _lib.ptr(x), n, _lib.ptr(f), _lib.ptr(g), _lib.stream_ptr()),
                       "mobo_rff_eval")
        return f, g

    def all_layers(self, x):
        """(L, n) device tensor: the sample of every layer of the chain at the rows of x."""
        return self._run(torch.as_tensor(x), False)[0]

    def value_and_grad(self, x):
        """(f (n,), df/dx (n, d)) of the top layer, batched (the reference's gradient path takes one point)."""
        f, g = self._run(torch.as_tensor(x), True)
        return f[-1], g

    def __call__(self, x, gradient=False):
        """``wrapper(x, gradient=False)`` of the reference: x (n, d) or (d,), numpy or torch; returns the values (n,)
        or, with gradient=True (one point, like the reference), d f / d x (d,).  numpy in -> numpy out."""
        is_np = isinstance(x, np.ndarray)
        xt = torch.as_tensor(x)
        if xt.ndim == 1:
            xt = xt[None, :]
        if gradient:
            assert xt.shape[0] == 1, "the gradient is defined for one point at a time (reference behaviour)"
            out = self._run(xt, True)[1][0]
        else:
            out = self._run(xt, False)[0][-1]
        return out.cpu().numpy() if is_np else out


def _np_to(dev, a):
    return torch.as_tensor(np.asarray(a, dtype=np.float64), device=dev)


def _layer_state(layer):
    vd = layer.variational_strategy._variational_distribution
    m = vd.variational_mean.detach().double()
    Lq = torch.tril(vd.chol_variational_covar.detach().double())
    return m, Lq @ Lq.T


def sample_layer(layer, input_dim, last=None, nFeatures=500, prior=False):
    """One function sample of `layer` on top of the sample `last` of the layer below (None for layer 0).  Draw order and
    distributions follow the reference line by line (posterior :311-338 / :364-404, prior :340-362 / :446-470)."""
    rs = np.random
    nF = int(nFeatures)
    dev = layer.variational_strategy._variational_distribution.variational_mean.device
    if dev.type != "cuda":
        raise RuntimeError("mobocmf_b200: move the model to a CUDA device before sampling functions (no CPU fallback)")
    d = int(input_dim)
    if layer.num_layer == 0:
        assert last is None
        if prior:
            lengthscale, alpha = 0.25 * d, 1.0
        else:
            th = layer.theta().detach().double().cpu().numpy()
            alpha, lengthscale = float(th[0]), th[1:]
        W = rs.normal(size=(nF, d)) / lengthscale
        b = rs.uniform(low=0.0, high=2 * np.pi, size=(nF, 1))
        Wt, bt = _np_to(dev, W), _np_to(dev, b)
        if prior:
            theta = _np_to(dev, rs.normal(loc=0.0, scale=1.0, size=nF))
        else:
            m, S = _layer_state(layer)
            Phi = _phi(layer._Zx().detach().double(), Wt, bt, alpha, nF)
            randomness = _np_to(dev, rs.normal(loc=0.0, scale=1.0, size=Phi.shape[0]))
            theta = posterior_weights(m, S, Phi, randomness)
        chain = [dict(kind=0, nF=nF, W=Wt, b=bt, theta=theta, alpha=float(alpha))]
        return RFFSample(chain, d, nF, dev)

    assert last is not None
    if prior:
        l1, lf, l2 = 10 * 0.25 * d, 1.0, 0.25 * d
        a1, af, a2, nu = 1.0, 1.0, 0.01, 1.0
    else:
        th = layer.theta().detach().double().cpu().numpy()     # [a1, v, af, lf, a2, l1.., l2..]
        a1, nu, af, lf, a2 = (float(v) for v in th[:5])
        l1, l2 = th[5:5 + d], th[5 + d:5 + 2 * d]
    W_x1 = rs.normal(size=(nF, d)) / l1
    W_f = rs.normal(size=nF) / lf
    W_x2 = rs.normal(size=(nF, d)) / l2
    b_x1 = rs.uniform(low=0.0, high=2 * np.pi, size=(nF, 1))
    b_x2 = rs.uniform(low=0.0, high=2 * np.pi, size=(nF, 1))
    t = {k: _np_to(dev, v) for k, v in dict(W_x1=W_x1, W_f=W_f, W_x2=W_x2, b_x1=b_x1, b_x2=b_x2).items()}
    a1f = a1 * af
    if prior:
        theta = _np_to(dev, rs.normal(loc=0.0, scale=1.0, size=3 * nF))
    else:
        m, S = _layer_state(layer)
        Zx = layer._Zx().detach().double()
        zf = layer._propagated_inducing_column().detach().double()
        Zxf = torch.cat([Zx, zf[:, None]], dim=1)
        W_x1f = torch.cat([t["W_x1"], t["W_f"][:, None]], dim=1)
        Phi = torch.cat([_phi(Zx, t["W_x1"], t["b_x1"], a1, nF) * zf * math.sqrt(nu),
                         _phi(Zxf, W_x1f, t["b_x1"], a1f, nF),
                         _phi(Zx, t["W_x2"], t["b_x2"], a2, nF)])
        randomness = _np_to(dev, rs.normal(loc=0.0, scale=1.0, size=Phi.shape[0]))
        theta = posterior_weights(m, S, Phi, randomness)
    top = dict(kind=1, nF=nF, theta=theta, alpha_x1=a1, alpha_x1f=a1f, alpha_x2=a2, nu_lin=nu, **t)
    return RFFSample(list(last.chain) + [top], d, nF, dev)
