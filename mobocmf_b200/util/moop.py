"""Multi-objective grid search over function samples — mirror of ``mobocmf/util/moop.py`` (class ``MOOP``, same
constructor / method names and return values) with the data-parallel parts on the GPU:

* every objective / constraint sample is evaluated on the whole candidate grid (``input_dim * grid_size`` uniform
  points + the observed inputs) by ONE ``mobo_rff_eval`` launch (``mobocmf_b200.rff.RFFSample``);
* the non-dominated cull (``compute_pareto_front`` / ``obtain_indices_pareto``, :141-185) is ``mobo_pareto_mask``;
* feasibility masks and the greedy min-max summary (:187-219) are device tensor ops over those results.
The per-objective constrained polish (``optimize_obj_globally``, :72-139) stays scipy SLSQP on the host, one point at
a time, exactly as in the reference: it is sequential by construction; its function / gradient calls go through the
same kernel.  The uniform grid is drawn from numpy's global generator like the reference (same seed, same grid)."""
import ctypes

import numpy as np
import scipy.optimize as spo
import torch

from .. import _lib


class NotFeasiblePoints(ValueError):
    pass


def _device_of(samples):
    for s in samples:
        dev = getattr(s, "device", None)
        if dev is not None:
            return torch.device(dev)
    return torch.device("cuda")


def _eval(sample, grid_dev):
    """Values of a sample on a device grid -> device (n,) tensor.  RFFSample: one kernel launch; any other callable
    (numpy in / out, like the reference's closures) is evaluated on the host and copied."""
    if hasattr(sample, "all_layers"):
        return sample(grid_dev)
    return torch.as_tensor(np.asarray(sample(grid_dev.cpu().numpy()), dtype=np.float64), device=grid_dev.device)


def pareto_mask(pts):
    """bool mask (device) of the non-dominated rows of pts (n, k) on a CUDA device (minimisation)."""
    pts = pts.to(dtype=torch.float64).contiguous()
    if not pts.is_cuda:
        raise RuntimeError("mobocmf_b200: pareto_mask needs a CUDA tensor (no CPU fallback)")
    n, k = pts.shape
    mask = torch.empty(n, dtype=torch.uint8, device=pts.device)
    with torch.cuda.device(pts.device):
        _lib.check(_lib.load().mobo_pareto_mask(_lib.ptr(pts), n, k, ctypes.c_void_p(mask.data_ptr()),
                                                _lib.stream_ptr()), "mobo_pareto_mask")
    return mask.bool()


class MOOP():

    def __init__(self, samples_objs, samples_cons, input_dim, grid_size=1000, pareto_set_size=None,
                 feasible_values=0.0, min_distance_between_points=1e-6):
        self.samples_objs = samples_objs
        self.samples_cons = samples_cons
        self.input_dim = input_dim
        self.bounds = [(0.0, 1.0)] * self.input_dim
        self.grid_size = grid_size
        self.pareto_set_size = pareto_set_size
        self.min_distance_between_points = min_distance_between_points
        self.feasible_values = feasible_values
        self.device = _device_of(list(samples_objs) + list(samples_cons))

    # ---- feasibility (moop.py:38-70) ----
    def find_feasible_grid(self, constraints, grid, feasible_values=0.0, allow_negative_constraints=False):
        """grid: (n, d) numpy array or tensor; returns the feasible rows as a DEVICE tensor, or None."""
        g = torch.as_tensor(grid, dtype=torch.float64).to(self.device)
        if not isinstance(feasible_values, np.ndarray):
            feasible_values = np.ones(max(self.input_dim, len(constraints))) * feasible_values
        vals = [_eval(c, g) for c in constraints]
        ok = torch.ones(g.shape[0], dtype=torch.bool, device=self.device)
        for i, v in enumerate(vals):
            ok &= v >= float(feasible_values[i])
        if bool(ok.any()):
            return g[ok]
        if not allow_negative_constraints:
            return None
        viol = torch.zeros(g.shape[0], dtype=torch.float64, device=self.device)
        for i, v in enumerate(vals):
            viol += torch.clamp(v - float(feasible_values[i]), max=0.0)
        return g[viol == viol[viol != 0].max()]

    # ---- constrained polish of one objective (moop.py:72-139); host SLSQP, kernel-evaluated samples ----
    def optimize_obj_globally(self, obj, cons, obj_evals, feasible_grid, constraint_tol=1e-6):
        assert self.input_dim == feasible_grid.shape[1]
        num_con = len(cons)
        fv = self.feasible_values
        if not isinstance(fv, np.ndarray):
            fv = np.ones(max(self.input_dim, num_con)) * fv
        best = int(torch.argmin(obj_evals))
        best_value = float(obj_evals[best])
        x0 = feasible_grid[best].detach().cpu().numpy().astype(np.float64)

        def val(fn, x):
            return float(np.asarray(fn(np.asarray(x, dtype=np.float64), gradient=False)).reshape(-1)[0])

        def grad(fn, x):
            return np.asarray(fn(np.asarray(x, dtype=np.float64), gradient=True), dtype=np.float64).reshape(-1)

        f = lambda x: val(obj, x)
        f_prime = lambda x: grad(obj, x)
        g_prime = lambda x: np.stack([grad(c, x) for c in cons]) if num_con else np.zeros((0, self.input_dim))

        for tol in (0.0, constraint_tol):
            g = lambda x, tol=tol: np.array([val(c, x) - tol - fv[i] for i, c in enumerate(cons)])
            opt_x = spo.fmin_slsqp(f, x0.copy(), bounds=self.bounds, disp=0, fprime=f_prime, f_ieqcons=g,
                                   fprime_ieqcons=g_prime)
            opt_x = np.clip(opt_x, a_min=0.0, a_max=1.0)
            if f(opt_x) < best_value and np.all(g(opt_x) >= -tol):
                return opt_x[None]
        return None

    # ---- non-dominated set (moop.py:141-185) ----
    @classmethod
    def compute_pareto_front(cls, pts):
        return pareto_mask(torch.as_tensor(pts))

    def obtain_indices_pareto(self, pts):
        return pareto_mask(torch.as_tensor(pts).to(self.device))

    # ---- greedy min-max summary in objective space (moop.py:187-219) ----
    def compute_pareto_front_and_set_summary_y_space(self, pareto_set, pareto_front, pareto_set_size):
        assert pareto_set_size > 0
        if pareto_set.shape[0] <= pareto_set_size:
            return pareto_set, pareto_front
        front = torch.as_tensor(pareto_front).to(self.device)
        dist = torch.sqrt(((front[:, None, :] - front[None, :, :]) ** 2).sum(-1))
        k = front.shape[1]
        subset = torch.zeros(pareto_set_size, dtype=torch.long, device=self.device)
        for i in range(k):
            subset[i] = torch.argmin(front[:, i])
        min_d = dist[subset[:k]].min(dim=0).values
        for c in range(k, pareto_set_size):
            subset[c] = torch.argmax(min_d)
            min_d = torch.minimum(min_d, dist[subset[c]])
        subset = subset.to(torch.as_tensor(pareto_set).device)
        return pareto_set[subset], pareto_front[subset.to(torch.as_tensor(pareto_front).device)]

    # ---- the whole search (moop.py:221-286) ----
    def compute_pareto_solution_from_samples(self, inputs, allow_negative_constraints=False):
        inputs_np = inputs.detach().cpu().numpy() if torch.is_tensor(inputs) else np.asarray(inputs)
        grid = np.concatenate((np.random.uniform(size=(self.input_dim * self.grid_size, self.input_dim)), inputs_np))
        grid = self.find_feasible_grid(self.samples_cons, grid, feasible_values=self.feasible_values,
                                       allow_negative_constraints=allow_negative_constraints)
        if grid is None:
            return None
        evals = torch.stack([_eval(obj, grid) for obj in self.samples_objs], dim=1)          # (n, k) on the device
        extra = []
        for i, obj in enumerate(self.samples_objs):
            opt_x = self.optimize_obj_globally(obj, self.samples_cons, evals[:, i], grid)
            if opt_x is not None:
                ox = torch.as_tensor(opt_x, dtype=torch.float64, device=self.device)
                if float(torch.cdist(grid, ox).min()) > 1e-6:
                    extra.append(ox)
        if extra:
            ex = torch.cat(extra, dim=0)
            grid = torch.cat([grid, ex], dim=0)
            evals = torch.cat([evals, torch.stack([_eval(obj, ex) for obj in self.samples_objs], dim=1)], dim=0)
        mask = self.obtain_indices_pareto(evals)
        pareto_set, pareto_front = grid[mask], evals[mask]
        if self.pareto_set_size is not None:
            pareto_set, pareto_front = self.compute_pareto_front_and_set_summary_y_space(pareto_set, pareto_front,
                                                                                       self.pareto_set_size)
        return pareto_set.cpu(), pareto_front.cpu(), self.samples_objs, self.samples_cons
