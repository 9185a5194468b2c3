"""Minimal parameter containers with GPyTorch's module / parameter naming (GPyTorch itself is not a dependency).

Only what the MFDGP hot path needs: the kernels instantiated at ``mobocmf/layers/mfdgp_hidden_layer.py:43-47,70-88``
(``ScaleKernel(RBFKernel)``, ``LinearKernel``, their product / sum), ``CholeskyVariationalDistribution``, the
Gaussian likelihood with an ``Interval`` noise constraint (``mobocmf/models/mfdgp.py:116``) and the
``num_likelihood_samples`` setting.  The classes hold parameters and transforms; the arithmetic is done by the CUDA
kernels (``functional.py``), never here.  ``state_dict`` keys match GPyTorch's so checkpoints and
warm starts interoperate.
"""
import contextlib
import math

import torch
from torch import nn


def inv_softplus(x):
    return x + torch.log(-torch.expm1(-x))


def inv_sigmoid(x):
    return torch.log(x) - torch.log(1 - x)


class Interval(nn.Module):
    """gpytorch.constraints.Interval: value = lower + (upper - lower) * sigmoid(raw); bounds kept as buffers
    (float32-rounded like upstream)."""

    def __init__(self, lower_bound, upper_bound):
        super().__init__()
        self.register_buffer("lower_bound", torch.as_tensor(lower_bound).float())
        self.register_buffer("upper_bound", torch.as_tensor(upper_bound).float())

    def transform(self, raw):
        return self.lower_bound + (self.upper_bound - self.lower_bound) * torch.sigmoid(raw)

    def inverse_transform(self, value):
        return inv_sigmoid((value - self.lower_bound) / (self.upper_bound - self.lower_bound))


class Positive(nn.Module):
    """gpytorch.constraints.Positive: value = softplus(raw)."""

    def __init__(self):
        super().__init__()
        self.register_buffer("lower_bound", torch.as_tensor(0.0))
        self.register_buffer("upper_bound", torch.as_tensor(math.inf))

    def transform(self, raw):
        return torch.nn.functional.softplus(raw)

    def inverse_transform(self, value):
        return inv_softplus(value)


class RBFKernel(nn.Module):
    def __init__(self, ard_num_dims, active_dims):
        super().__init__()
        self.ard_num_dims = ard_num_dims
        self.active_dims = tuple(active_dims)
        self.raw_lengthscale = nn.Parameter(torch.zeros(1, ard_num_dims))
        self.raw_lengthscale_constraint = Positive()

    @property
    def lengthscale(self):
        return self.raw_lengthscale_constraint.transform(self.raw_lengthscale)

    def initialize(self, lengthscale):
        value = torch.as_tensor(lengthscale).to(self.raw_lengthscale)
        self.raw_lengthscale.data.copy_(self.raw_lengthscale_constraint.inverse_transform(value).expand_as(
            self.raw_lengthscale))
        return self


class ScaleKernel(nn.Module):
    def __init__(self, base_kernel):
        super().__init__()
        self.base_kernel = base_kernel
        self.raw_outputscale = nn.Parameter(torch.zeros(()))
        self.raw_outputscale_constraint = Positive()

    @property
    def outputscale(self):
        return self.raw_outputscale_constraint.transform(self.raw_outputscale)

    def initialize(self, outputscale):
        value = torch.as_tensor(outputscale).to(self.raw_outputscale)
        self.raw_outputscale.data.copy_(self.raw_outputscale_constraint.inverse_transform(value))
        return self


class LinearKernel(nn.Module):
    def __init__(self, active_dims):
        super().__init__()
        self.active_dims = tuple(active_dims)
        self.raw_variance = nn.Parameter(torch.zeros(1, 1))
        self.raw_variance_constraint = Positive()

    @property
    def variance(self):
        return self.raw_variance_constraint.transform(self.raw_variance)

    def initialize(self, variance):
        value = torch.as_tensor(variance).to(self.raw_variance)
        self.raw_variance.data.copy_(self.raw_variance_constraint.inverse_transform(value).expand_as(
            self.raw_variance))
        return self


class _CompositeKernel(nn.Module):
    def __init__(self, *kernels):
        super().__init__()
        self.kernels = nn.ModuleList(kernels)


class AdditiveKernel(_CompositeKernel):
    pass


class ProductKernel(_CompositeKernel):
    pass


class ZeroMean(nn.Module):
    pass


class CholeskyVariationalDistribution(nn.Module):
    """q(u) = N(variational_mean, L L^T), L = tril(chol_variational_covar) (the tril is applied by the kernels)."""

    def __init__(self, num_inducing_points):
        super().__init__()
        self.variational_mean = nn.Parameter(torch.zeros(num_inducing_points))
        self.chol_variational_covar = nn.Parameter(torch.eye(num_inducing_points))


class HomoskedasticNoise(nn.Module):
    def __init__(self, noise_constraint):
        super().__init__()
        self.raw_noise = nn.Parameter(torch.zeros(1))
        self.raw_noise_constraint = noise_constraint

    @property
    def noise(self):
        return self.raw_noise_constraint.transform(self.raw_noise)


class GaussianLikelihood(nn.Module):
    """gpytorch.likelihoods.GaussianLikelihood with a homoskedastic noise (mobocmf/models/mfdgp.py:116-121)."""

    def __init__(self, noise_constraint):
        super().__init__()
        self.noise_covar = HomoskedasticNoise(noise_constraint)

    @property
    def raw_noise(self):
        return self.noise_covar.raw_noise

    @property
    def noise(self):
        return self.noise_covar.noise

    @noise.setter
    def noise(self, value):
        nc = self.noise_covar
        value = torch.as_tensor(value).to(nc.raw_noise)
        nc.raw_noise.data.copy_(nc.raw_noise_constraint.inverse_transform(value).expand_as(nc.raw_noise))

    def expected_log_prob(self, target, dist):
        """E_q[log N(y | f, noise)] per point (upstream GaussianLikelihood.expected_log_prob)."""
        noise = self.noise.reshape(())
        res = ((target - dist.mean).square() + dist.variance) / noise + noise.log() + math.log(2 * math.pi)
        return res.mul(-0.5)

    def __call__(self, dist):
        """Marginal: adds the noise to the (raw) variance."""
        return GaussianMoments(dist.mean, dist.raw_variance + self.noise.reshape(()))


MIN_VARIANCE = 1e-10   # gpytorch.settings.min_variance, fp64 (quirk Q9)


class GaussianMoments(object):
    """Stand-in for gpytorch.distributions.MultivariateNormal where only the marginal moments are consumed:
    ``.mean`` and ``.variance`` (floored at 1e-10 on every read, like upstream); ``.raw_variance`` is unfloored."""

    def __init__(self, mean, raw_variance):
        self._mean = mean
        self.raw_variance = raw_variance

    @property
    def mean(self):
        return self._mean

    @property
    def loc(self):
        return self._mean

    @property
    def variance(self):
        return self.raw_variance.clamp_min(MIN_VARIANCE)

    @property
    def stddev(self):
        return self.variance.sqrt()


class _Settings(object):
    """gpytorch.settings subset: ``num_likelihood_samples`` context manager (the reference always uses 1)."""
    _num_likelihood_samples = 10

    class num_likelihood_samples(contextlib.ContextDecorator):
        def __init__(self, value):
            self.value = value

        def __enter__(self):
            self.prev = _Settings._num_likelihood_samples
            _Settings._num_likelihood_samples = self.value

        def __exit__(self, *a):
            _Settings._num_likelihood_samples = self.prev
            return False

        @staticmethod
        def value_():
            return _Settings._num_likelihood_samples


settings = _Settings
