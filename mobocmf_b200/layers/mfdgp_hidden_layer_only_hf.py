"""Only-highest-fidelity twin of the layer — mirror of ``mobocmf/layers/mfdgp_hidden_layer_only_hf.py``: layer >= 1
starts with k_lin.variance = 0, k_x1 / k_f outputscale = 0, k_x2 outputscale = 1 (:85-89) and freezes the k_x1, k_f,
k_lin parameters (:193-199).  Same kernels, different constants."""
from .mfdgp_hidden_layer import MFDGPHiddenLayer as _Base


class MFDGPHiddenLayer(_Base):
    only_hf = True

    def _initial_scales(self):
        return 0.0, 0.0, 1.0, 0.0      # a1, a_f, a2, v_lin

    def _freeze_after_init(self):
        if self.num_layer > 0:
            k_x1, k_sum = self.covar_module.kernels[0].kernels
            k_lin, k_f = k_sum.kernels
            k_x1.base_kernel.raw_lengthscale.requires_grad = False
            k_x1.raw_outputscale.requires_grad = False
            k_f.base_kernel.raw_lengthscale.requires_grad = False
            k_f.raw_outputscale.requires_grad = False
            k_lin.raw_variance.requires_grad = False
