"""One sparse-GP layer of the multi-fidelity deep GP — host-side mirror of
``mobocmf/layers/mfdgp_hidden_layer.py`` (class names, constructor signature, parameter tree, call semantics), with
the arithmetic routed into the sm_100a kernels of ``mobocmf_b200.functional``.

Differences from the reference that are deliberate:
* no GPyTorch: the module tree is rebuilt from ``mobocmf_b200.gp`` with GPyTorch's parameter names;
* the training-mode normals can be injected (``eps=``) because a GPU build cannot replay the reference's CPU
  generator stream; by default they are drawn on the device;
* the eval branch computes only the diagonal of the predictive covariance (the reference materialises R x R and
  reads its diagonal).
``sample_from_posterior`` / ``sample_from_prior`` (RFF function samples, :288-514) return ``mobocmf_b200.rff.RFFSample``
objects evaluated by ``mobo_rff_eval``.
"""
from typing import Optional

import torch
from torch import Tensor, nn

from .. import functional as F
from ..gp import (AdditiveKernel, CholeskyVariationalDistribution, GaussianMoments, LinearKernel, ProductKernel,
                  RBFKernel, ScaleKernel, ZeroMean, settings)


def _dense_kernel_for_init(covar_module, num_layer, Z):
    """K(Z, Z) in plain torch, used ONCE at construction for the q(u) initial covariance of the top layer
    (layers/mfdgp_hidden_layer.py:131-132).  Not on the hot path."""
    def rbf(k, X):
        Xs = X[:, list(k.base_kernel.active_dims)] / k.base_kernel.lengthscale
        d2 = (Xs[:, None, :] - Xs[None, :, :]).pow(2).sum(-1)
        return k.outputscale * torch.exp(-0.5 * d2)
    if num_layer == 0:
        return rbf(covar_module, Z)
    k_x1, k_sum = covar_module.kernels[0].kernels
    k_lin, k_f = k_sum.kernels
    k_x2 = covar_module.kernels[1]
    f = Z[:, list(k_lin.active_dims)]
    lin = (f * k_lin.variance.sqrt()) @ (f * k_lin.variance.sqrt()).T
    return rbf(k_x1, Z) * (lin + rbf(k_f, Z)) + rbf(k_x2, Z)


def _psd_safe_cholesky(A, jitter=1e-8, max_tries=3):
    L, info = torch.linalg.cholesky_ex(A)
    if not bool(info.any()):
        return L
    Ap = A.clone()
    prev = 0.0
    for i in range(max_tries):
        new = jitter * (10 ** i)
        Ap.diagonal().add_(new - prev)
        prev = new
        L, info = torch.linalg.cholesky_ex(Ap)
        if not bool(info.any()):
            return L
    raise RuntimeError("NotPSDError: initial variational covariance is not positive definite")


class UnwhitenedVariationalStrategy(nn.Module):
    """Parameter holder with GPyTorch's names: ``inducing_points`` / ``variational_params_initialized`` buffers and
    ``_variational_distribution``.  learn_inducing_locations is always False in the reference
    (layers/mfdgp_hidden_layer.py:142,146)."""

    def __init__(self, model, inducing_points, variational_distribution, learn_inducing_locations=False,
                 jitter_val=None):
        super().__init__()
        object.__setattr__(self, "model", model)
        if learn_inducing_locations:
            raise NotImplementedError("the reference never learns the inducing locations")
        self._buffers["inducing_points"] = inducing_points.clone()
        self.register_buffer("variational_params_initialized", torch.tensor(1))
        self._variational_distribution = variational_distribution
        # settings.variational_cholesky_jitter by dtype of the inducing points (quirk Q5)
        if jitter_val is None:
            jitter_val = 1e-6 if inducing_points.dtype == torch.float64 else 1e-4
        self.jitter_val = jitter_val

    @property
    def original_inducing_points(self):
        return self._buffers["inducing_points"]

    @property
    def inducing_points(self) -> Tensor:
        return self._buffers["inducing_points"]

    @property
    def variational_distribution(self):
        vd = self._variational_distribution
        L = torch.tril(vd.chol_variational_covar)
        return GaussianMoments(vd.variational_mean, (L ** 2).sum(-1))

    def kl_divergence(self):
        return self.model._kl_divergence()


class MFDGUnwhitenedVariationalStrategy(UnwhitenedVariationalStrategy):
    """Strategy whose inducing inputs are ``[Z, mean_{l-1}(Z)]`` (layers/mfdgp_hidden_layer.py:520-559).  The
    previous layer is registered as a sub-module exactly like the reference (quirk Q12)."""

    def __init__(self, model, inducing_points: Tensor, variational_distribution: CholeskyVariationalDistribution,
                 learn_inducing_locations: bool = True, jitter_val: Optional[float] = None,
                 previous_layer: Optional["MFDGPHiddenLayer"] = None):
        super().__init__(model, inducing_points, variational_distribution, learn_inducing_locations, jitter_val)
        self.previous_layer = previous_layer

    @property
    def inducing_points(self) -> Tensor:
        Zo = self._buffers["inducing_points"]
        if self.previous_layer is None:
            return Zo
        zx = Zo[:, :-1]
        return torch.cat((zx, self.model._propagated_inducing_column()[:, None]), 1)


class MFDGPHiddenLayer(nn.Module):
    only_hf = False

    def __init__(self, num_layer, input_dims, inducing_points, inducing_values, num_fidelities, init_lengthscale,
                 y_high_std=1.0, num_samples_for_acquisition=25, previously_trained_layer=None,
                 init_params_to_prior_and_fix_them=False, previous_layer_in_hierarchy=None):
        super().__init__()
        self.init_params_to_prior_and_fix_them = init_params_to_prior_and_fix_them
        self.num_layer = num_layer
        self.input_dims = input_dims
        num_inducing = inducing_points.shape[0]
        self.num_inducing = num_inducing
        self.output_dims = None

        if num_layer == 0:
            covar_module = ScaleKernel(RBFKernel(ard_num_dims=input_dims, active_dims=range(input_dims)))
            covar_module.base_kernel.initialize(lengthscale=init_lengthscale)
            covar_module.initialize(outputscale=1.0)
            if self.init_params_to_prior_and_fix_them:
                covar_module.base_kernel.initialize(lengthscale=0.25 * input_dims)
                covar_module.initialize(outputscale=1.0)
        else:
            D = list(range(input_dims))
            k_x_1 = ScaleKernel(RBFKernel(ard_num_dims=input_dims - 1, active_dims=D[0:input_dims - 1]))
            k_f = ScaleKernel(RBFKernel(ard_num_dims=1, active_dims=D[input_dims - 1:input_dims]))
            k_x_2 = ScaleKernel(RBFKernel(ard_num_dims=input_dims - 1, active_dims=D[0:input_dims - 1]))
            k_lin = LinearKernel(active_dims=D[input_dims - 1:input_dims])
            k_x_1.base_kernel.initialize(lengthscale=init_lengthscale * 10.0)
            k_f.base_kernel.initialize(lengthscale=1.0)
            k_x_2.base_kernel.initialize(lengthscale=init_lengthscale)
            a1, af, a2, vl = self._initial_scales()
            k_lin.initialize(variance=torch.ones(1) * vl)
            k_x_1.initialize(outputscale=a1)
            k_f.initialize(outputscale=af)
            k_x_2.initialize(outputscale=a2)
            if self.init_params_to_prior_and_fix_them:
                k_x_1.base_kernel.initialize(lengthscale=10 * 0.25 * (input_dims - 1))
                k_f.base_kernel.initialize(lengthscale=1.0)
                k_x_2.base_kernel.initialize(lengthscale=0.25 * (input_dims - 1))
                k_lin.initialize(variance=torch.ones(1) * 1.0)
                k_x_1.initialize(outputscale=1.0)
                k_f.initialize(outputscale=1.0)
                k_x_2.initialize(outputscale=0.01)
            covar_module = AdditiveKernel(ProductKernel(k_x_1, AdditiveKernel(k_lin, k_f)), k_x_2)

        if previously_trained_layer is not None:
            covar_module.load_state_dict(previously_trained_layer.covar_module.state_dict())

        variational_distribution = CholeskyVariationalDistribution(num_inducing_points=num_inducing)
        with torch.no_grad():
            variational_distribution.variational_mean.copy_(inducing_values)
            if num_layer == num_fidelities - 1:
                cov = _dense_kernel_for_init(covar_module, num_layer, inducing_points) * (1e-2 * y_high_std ** 2) ** 2
                variational_distribution.chol_variational_covar.copy_(_psd_safe_cholesky(cov))
            else:
                variational_distribution.chol_variational_covar.copy_(
                    _psd_safe_cholesky(torch.eye(num_inducing) * 1e-8))

        if num_layer == 0:
            self.variational_strategy = UnwhitenedVariationalStrategy(self, inducing_points,
                                                                      variational_distribution,
                                                                      learn_inducing_locations=False)
        else:
            self.variational_strategy = MFDGUnwhitenedVariationalStrategy(
                self, inducing_points, variational_distribution, learn_inducing_locations=False,
                previous_layer=previous_layer_in_hierarchy)

        self.mean_module = ZeroMean()
        self.covar_module = covar_module

        if previously_trained_layer is not None:
            self.samples = torch.ones([num_samples_for_acquisition]) * previously_trained_layer.samples
        else:
            self.samples = torch.normal(mean=torch.zeros([num_samples_for_acquisition]),
                                        std=torch.ones([num_samples_for_acquisition]))[:, None]
        self.num_samples_for_acquisition = num_samples_for_acquisition
        self._eval_mode = False
        self._ops_cache = None
        self._dev_cache = {}

        if self.init_params_to_prior_and_fix_them:
            for p in self.covar_module.parameters():
                p.requires_grad = False
        self._freeze_after_init()

    # ---- hooks for the only-highest-fidelity twin (layers/mfdgp_hidden_layer_only_hf.py:85-89,193-199) ----
    def _initial_scales(self):
        return 1.0, 1.0, 0.01, 1.0     # a1, a_f, a2, v_lin  (layers/mfdgp_hidden_layer.py:84-88)

    def _freeze_after_init(self):
        pass

    # ---- state that must not be pickled / deep-copied (copy_uncond, util/blackbox_mfdgp_fitter.py:383) ----
    def __getstate__(self):
        state = self.__dict__.copy()
        state["_ops_cache"] = None
        state["_dev_cache"] = {}
        return state

    def print_lengthscales_and_outputscale(self, custom_print):
        """layers/mfdgp_hidden_layer.py:190-224: the constrained hyper-parameters, named as in the reference."""
        th = self.theta().detach().cpu().numpy()
        if self.num_layer == 0:
            custom_print({"l0_lengthscale:": th[1:], "l0_outputscale:": float(th[0])})
            return
        d = self.x_dims
        custom_print({"l1_lengthscale_x1:": th[5:5 + d], "l1_lengthscale_f:": th[3:4], "l1_lengthscale_x2:": th[5 + d:],
                      "l1_alpha_x1:": float(th[0]), "l1_alpha_f:": float(th[2]), "l1_alpha_x1f:": float(th[0] * th[2]),
                      "l1_alpha_x2:": float(th[4]), "l1_nu_lin:": float(th[1])})

    def train_mode(self):
        self._eval_mode = False

    def eval_mode(self):
        self._eval_mode = True

    # ---- random-Fourier-feature function samples (layers/mfdgp_hidden_layer.py:311-514) ----
    def sample_from_posterior(self, input_dim, sample_from_posterior_last_layer=None, nFeatures=500):
        from .. import rff
        return rff.sample_layer(self, input_dim, sample_from_posterior_last_layer, nFeatures, prior=False)

    def sample_from_prior(self, input_dim, sample_from_prior_last_layer=None, nFeatures=500):
        from .. import rff
        return rff.sample_layer(self, input_dim, sample_from_prior_last_layer, nFeatures, prior=True)

    # ---- kernel-facing views of the parameters ----
    @property
    def kind(self):
        return 0 if self.num_layer == 0 else 1

    @property
    def x_dims(self):
        return self.input_dims if self.num_layer == 0 else self.input_dims - 1

    def theta(self):
        """Constrained hyper-parameters in the kernels' layout (include/mobocmf_b200.h)."""
        cm = self.covar_module
        if self.num_layer == 0:
            return torch.cat([cm.outputscale.reshape(1), cm.base_kernel.lengthscale.reshape(-1)])
        k_x1, k_sum = cm.kernels[0].kernels
        k_lin, k_f = k_sum.kernels
        k_x2 = cm.kernels[1]
        return torch.cat([k_x1.outputscale.reshape(1), k_lin.variance.reshape(1), k_f.outputscale.reshape(1),
                          k_f.base_kernel.lengthscale.reshape(1), k_x2.outputscale.reshape(1),
                          k_x1.base_kernel.lengthscale.reshape(-1), k_x2.base_kernel.lengthscale.reshape(-1)])

    def _Zx(self):
        Zo = self.variational_strategy.original_inducing_points
        key = ("Zx", Zo.data_ptr(), Zo._version)
        if self._dev_cache.get("Zx_key") != key:
            self._dev_cache["Zx"] = Zo[:, :self.x_dims].contiguous()
            self._dev_cache["Zx_key"] = key
        return self._dev_cache["Zx"]

    def _samples_on(self, device):
        key = ("samples", str(device))
        if key not in self._dev_cache:
            self._dev_cache[key] = self.samples.to(device=device, dtype=torch.float64).reshape(-1).contiguous()
        return self._dev_cache[key]

    def _propagated_inducing_column(self):
        """Last column of this layer's inducing inputs: previous_layer(Z[:, :d]).mean
        (layers/mfdgp_hidden_layer.py:556-557).  With shared Z the previous layer hits the x == Z shortcut of the
        upstream strategy and returns its variational mean exactly (quirk Q4)."""
        prev = self.variational_strategy.previous_layer
        Zx = self._Zx()
        shared = self._dev_cache.get("shared")
        if shared is None:
            pz = prev._Zx()
            shared = pz.shape == Zx.shape and bool(torch.equal(pz, Zx))
            self._dev_cache["shared"] = shared
        if shared:
            return prev.variational_strategy._variational_distribution.variational_mean
        if prev.num_layer != 0:
            raise NotImplementedError("non-shared inducing inputs above layer 1 crash in the reference too (F3)")
        with settings.num_likelihood_samples(1):
            mu, _ = prev._moments(Zx)
        return mu

    def _op_inputs(self):
        vd = self.variational_strategy._variational_distribution
        return list(self.covar_module.parameters()) + [vd.variational_mean, vd.chol_variational_covar]

    def _invalidate(self, *_):
        self._ops_cache = None

    def operators(self):
        """Per-step operator buffer (L_p, L_p^-1, L_p^-1 L_q, ..., KL).  Computed once per parameter version:
        in training mode with autograd (one M^3 chain per step, shared by every forward of the step and by the
        KL term); in eval mode under no_grad (parameters are constants for the acquisition, only dX flows)."""
        grad_mode = torch.is_grad_enabled() and self.training
        ins = self._op_inputs()
        prev = getattr(self.variational_strategy, "previous_layer", None)
        if prev is not None:
            ins = ins + [prev.variational_strategy._variational_distribution.variational_mean]
        key = (grad_mode, tuple((p.data_ptr(), p._version) for p in ins))
        if self._ops_cache is not None and self._ops_cache[0] == key:
            return self._ops_cache[1]
        with torch.set_grad_enabled(grad_mode):
            vd = self.variational_strategy._variational_distribution
            zf = self._propagated_inducing_column() if self.num_layer > 0 else None
            theta = self.theta()
            ops = F.layer_operators(theta, zf, vd.variational_mean, vd.chol_variational_covar, self._Zx(),
                                    self.kind, self.variational_strategy.jitter_val)
        if ops.requires_grad:
            ops.register_hook(self._invalidate)
        self._ops_cache = (key, (ops, theta, zf))
        return self._ops_cache[1]

    def _kl_divergence(self):
        ops, _, _ = self.operators()
        return F.ops_kl(ops, self.num_inducing)

    def _moments(self, x, mu_prev=None, var_prev=None, eps=None, f_direct=None, xrep=1, prep=1, eps_mod=None,
                 R=None):
        """(mean, raw variance) of q(f_l) for R rows through the fused row kernel."""
        ops, theta, zf = self.operators()
        if not (torch.is_grad_enabled() and self.training):
            theta, zf = theta.detach(), (None if zf is None else zf.detach())
        return F.layer_rows(ops, theta, zf, self._Zx(), x, mu_prev, var_prev, eps, f_direct, kind=self.kind,
                            xrep=xrep, prep=prep, eps_mod=eps_mod, R=R, training=self.training)

    def _equals_inducing(self, x):
        """x equals this layer's inducing inputs row for row (the upstream `torch.equal(x, Z)` shortcut, quirk Q4).
        A device comparison costs a synchronisation; callers that know the answer on the host say so: the fitter's
        loader attaches a host copy (`x._mobo_host`), the graph-captured conditioned step marks its static buffers
        (`x._mobo_not_z`) after checking every minibatch on the host."""
        Zx = self._Zx()
        if x.shape != Zx.shape or getattr(x, "_mobo_not_z", False):
            return False
        host = getattr(x, "_mobo_host", None)
        if host is not None:
            key = ("Zx_host", Zx.data_ptr(), Zx._version)
            if self._dev_cache.get("Zx_host_key") != key:
                self._dev_cache["Zx_host"], self._dev_cache["Zx_host_key"] = Zx.detach().cpu(), key
            return bool(torch.equal(host, self._dev_cache["Zx_host"]))
        return bool(torch.equal(x, Zx))

    def forward(self, x):
        raise RuntimeError("MFDGPHiddenLayer.forward (the lazy prior over cat[Z, X]) has no dense counterpart here; "
                           "call the layer")

    def __call__(self, x, *other_inputs, eps=None, **kwargs):
        """layers/mfdgp_hidden_layer.py:245-286.  Returns the q(f) moments: shape (1, R) for deterministic inputs
        (layer 0, leading num_likelihood_samples dim, quirk Q3), (R,) otherwise."""
        if len(other_inputs) == 0:
            assert x.shape[-1] == self.input_dims
            if self.num_layer == 0 and self._equals_inducing(x):
                q = self.variational_strategy.variational_distribution        # quirk Q4 shortcut: N(m, S)
                mu, var = q.mean, q.raw_variance
            else:
                mu, var = self._moments(x.contiguous())
            ns = settings.num_likelihood_samples.value_()
            return GaussianMoments(mu.unsqueeze(0).expand(ns, -1), var.unsqueeze(0).expand(ns, -1))
        if isinstance(x, GaussianMoments):
            raise ValueError
        if len(other_inputs) != 1:
            raise ValueError("one propagated input per layer")
        inp = other_inputs[0]
        R = x.shape[0]
        assert x.shape[-1] + 1 == self.input_dims
        if isinstance(inp, GaussianMoments):
            mu_p = inp.mean.reshape(-1)
            var_p = inp.raw_variance.reshape(-1)
            if self._eval_mode:
                S = self.num_samples_for_acquisition
                mu, var = self._moments(x.contiguous(), mu_p.contiguous(), var_p.contiguous(),
                                        self._samples_on(x.device), eps_mod=S, R=R)
            else:
                if eps is None:
                    eps = torch.randn(R, device=x.device, dtype=torch.float32)
                eps = eps.to(device=x.device, dtype=torch.float64).reshape(-1).contiguous()
                mu, var = self._moments(x.contiguous(), mu_p.contiguous(), var_p.contiguous(), eps, eps_mod=R, R=R)
        else:
            mu, var = self._moments(x.contiguous(), f_direct=inp.T.reshape(-1).contiguous().to(x.dtype), R=R)
        return GaussianMoments(mu, var)
