"""Small helpers mirrored from ``mobocmf/util/util.py`` (the two used by the lengthscale heuristic, :27-33, and the
seeding helper :70-72)."""
import numpy as np
import torch


def triu_indices(n, offset=0):
    rows, cols = torch.triu_indices(n, n, offset=offset)
    return torch.stack((rows, cols), dim=0)


def compute_dist(x):
    return torch.sum(x ** 2, 1, keepdims=True) - 2.0 * x.mm(x.T) + torch.sum(x ** 2, 1, keepdims=True).T


def reset_random_state(seed):
    torch.manual_seed(seed)
    np.random.seed(seed)
