"""ELBO of the multi-fidelity deep GP — mirror of ``mobocmf/mlls/variational_elbo_mf.py``."""
import torch


class VariationalELBOMF(object):

    def __init__(self, model, num_data, num_fidelities):
        self.likelihood = None
        self.model = model
        self.num_data = num_data
        self.num_fidelities = num_fidelities

    def forward(self, l_approximate_dist_f, target, fidelities, include_kl_term=True):
        assert target.shape[0] <= target.shape[1]   # the target must be (1, B)
        num_batch = target.shape[1]
        data_term = 0.0
        for i in range(self.num_fidelities):
            mask = fidelities.T == i
            if mask.sum() != 0:
                likelihood = getattr(self.model, self.model.name_hidden_layer_likelihood + str(i))
                data_term = data_term + likelihood.expected_log_prob(target, l_approximate_dist_f[i])[mask].sum()
        if include_kl_term is False:
            return data_term
        kl_divergence = self.model.variational_strategy.kl_divergence()
        return data_term - kl_divergence * num_batch / self.num_data, kl_divergence * num_batch / self.num_data

    __call__ = forward
