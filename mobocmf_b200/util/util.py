"""Small host helpers mirrored from ``mobocmf/util/util.py``: the two used by the lengthscale heuristic (:27-33), the
seeding helper (:70-72), the output preprocessing (:36-67: identity scaling, kept for call compatibility) and the
pickle helpers (:7-25; ``dill`` when it is installed, the standard ``pickle`` otherwise)."""
import os
import pickle

import numpy as np
import torch

try:
    import dill as _pickler
except ImportError:      # dill is not a dependency of the GPU path
    _pickler = pickle


def create_path(folder):
    os.makedirs(folder, exist_ok=True)


def save_pickle(folder, filename, content):
    create_path(folder)
    with open(os.path.join(folder, filename), "wb") as fw:
        _pickler.dump(content, fw)


def read_pickle(folder, filename):
    with open(os.path.join(folder, filename), "rb") as fr:
        return _pickler.load(fr)


def preprocess_outputs(*args):
    """Outputs as double tensors followed by (mean, std) = (0.0, 1.0): the reference deliberately does not standardise
    (the linear dependencies between fidelities would break)."""
    y_mean, y_std = 0.0, 1.0
    return [torch.from_numpy((y - y_mean) / y_std).double() for y in args] + [y_mean, y_std]


def preprocess_outputs_two_fidelities(y_low, y_high):
    y_mean, y_std = 0.0, 1.0
    return (torch.from_numpy((y_low - y_mean) / y_std).double(), torch.from_numpy((y_high - y_mean) / y_std).double(),
            y_mean, y_std)


def triu_indices(n, offset=0):
    rows, cols = torch.triu_indices(n, n, offset=offset)
    return torch.stack((rows, cols), dim=0)


def compute_dist(x):
    return torch.sum(x ** 2, 1, keepdims=True) - 2.0 * x.mm(x.T) + torch.sum(x ** 2, 1, keepdims=True).T


def reset_random_state(seed):
    torch.manual_seed(seed)
    np.random.seed(seed)
