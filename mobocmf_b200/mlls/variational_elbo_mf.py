"""ELBO of the multi-fidelity deep GP — mirror of ``mobocmf/mlls/variational_elbo_mf.py``."""
import torch


class VariationalELBOMF(object):

    def __init__(self, model, num_data, num_fidelities):
        self.likelihood = None
        self.model = model
        self.num_data = num_data
        self.num_fidelities = num_fidelities

    def __getstate__(self):
        # the fused-step binding (ctypes descriptors, device workspace, captured graphs) is rebuilt on demand and must
        # not travel through copy.deepcopy / pickle (copy_uncond, mobocmf/util/blackbox_mfdgp_fitter.py:383)
        state = self.__dict__.copy()
        state.pop("_fused_step", None)
        return state

    def forward(self, l_approximate_dist_f, target, fidelities, include_kl_term=True):
        assert target.shape[0] <= target.shape[1]   # the target must be (1, B)
        num_batch = target.shape[1]
        data_term = 0.0
        for i in range(self.num_fidelities):
            # the reference tests `mask.sum() != 0` and sums `ell[mask]` (variational_elbo_mf.py:33-38): both force a
            # device synchronisation (and a data-dependent shape); the masked sum below is the same number, adds an
            # exact zero for a fidelity without points, and can be captured in a CUDA graph
            mask = fidelities.T == i
            likelihood = getattr(self.model, self.model.name_hidden_layer_likelihood + str(i))
            dist = l_approximate_dist_f[i]
            S = getattr(dist, "samples_per_point", 1)
            if S > 1:   # S-sample extension: average the per-sample terms of each point
                ell = likelihood.expected_log_prob(target.reshape(-1, 1).expand(-1, S).reshape(-1), dist)
                ell = ell.reshape(-1, S).mean(1)[None, :]
            else:
                ell = likelihood.expected_log_prob(target, dist)
            data_term = data_term + torch.where(mask, ell, torch.zeros((), dtype=ell.dtype, device=ell.device)).sum()
        if include_kl_term is False:
            return data_term
        kl_divergence = self.model.variational_strategy.kl_divergence()
        return data_term - kl_divergence * num_batch / self.num_data, kl_divergence * num_batch / self.num_data

    __call__ = forward
