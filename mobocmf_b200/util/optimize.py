"""Stand-in for ``botorch.optim.optimize_acqf`` as called at
``mobocmf/acquisition_functions/JESMOC_MFDGP.py:142-143,159-160`` (q = 1): quasi-random raw samples evaluated in one
batch under no_grad, the best ``num_restarts`` refined jointly by L-BFGS-B (scipy) with autograd gradients.

What makes it a GPU workload (each L-BFGS iteration evaluates 2 K model chains on a handful of rows: pure launch
latency, ~600 kernel launches for three fidelities x six black boxes):

* ``optimize_acqf_multi`` optimises SEVERAL acquisition functions at once - the per-fidelity loop of
  ``JESMOC_MFDGP.get_nextpoint_coupled`` (reference lines 151-168) becomes one L-BFGS-B run over all restarts of all
  fidelities (the objective is separable, so the optima are those of the separate runs), i.e. one batch per iteration;
* on CUDA the value-and-gradient evaluation of that batch (forward of every model chain, sum, backward to dX) is
  captured ONCE in a CUDA graph and replayed per iteration: one graph launch, one small H2D copy of the iterate and one
  D2H copy of (loss, gradient) instead of hundreds of Python-driven launches.  The model parameters are constants
  during the optimisation (eval-mode operator caches), so the graph is valid for the whole run; it is rebuilt per call.
"""
import numpy as np
import torch
from scipy.optimize import minimize


class GraphedValueAndGrad(object):
    """X (n, 1, d) -> (-sum fn(X), d(-sum) / dX, fn(X)) replayed from a CUDA graph.  ``fn`` must be capturable: no host
    synchronisation, shapes fixed by X's (the MFDGP acquisition chain is, once its eval-mode operators are cached)."""

    def __init__(self, fn, n, d, device, warmup=1):
        self.n, self.d = n, d
        self.X = torch.zeros(n, 1, d, dtype=torch.float64, device=device, requires_grad=True)
        self.X._mobo_not_z = True        # never the inducing inputs: keeps the (synchronising) shortcut test out of capture
        self.out = torch.zeros(1 + n * d + n, dtype=torch.float64, device=device)       # [loss | grad | values]
        self.host_in = torch.zeros(n, 1, d, dtype=torch.float64).pin_memory()
        self.host_out = torch.zeros(1 + n * d + n, dtype=torch.float64).pin_memory()

        def body():
            v = fn(self.X).double()
            loss = -v.sum()
            g, = torch.autograd.grad(loss, self.X)
            self.out[0:1].copy_(loss.detach().reshape(1))
            self.out[1:1 + n * d].copy_(g.reshape(-1))
            self.out[1 + n * d:].copy_(v.detach().reshape(-1))

        side = torch.cuda.Stream(device=device)
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                body()
        torch.cuda.current_stream(device).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            body()
        self.replays = 0

    def __call__(self, flat):
        self.host_in.copy_(torch.from_numpy(np.ascontiguousarray(flat, dtype=np.float64)).reshape(self.n, 1, self.d))
        with torch.no_grad():
            self.X.copy_(self.host_in, non_blocking=True)
        self.graph.replay()
        self.host_out.copy_(self.out, non_blocking=True)
        torch.cuda.current_stream(self.X.device).synchronize()
        self.replays += 1
        o = self.host_out.numpy()
        n, d = self.n, self.d
        return float(o[0]), o[1:1 + n * d].copy(), o[1 + n * d:].copy()


def optimize_acqf_multi(acq_functions, bounds, q=1, num_restarts=5, raw_samples=200, options=None, seed=None,
                        use_cuda_graph=None, return_info=False):
    """``optimize_acqf`` for every function of ``acq_functions`` in ONE multi-start L-BFGS-B run.  Returns
    [(candidate (1, d), value)] in the order of the functions (and an info dict with ``return_info``)."""
    assert q == 1
    options = options or {}
    bounds = torch.as_tensor(bounds, dtype=torch.double)
    dev = bounds.device
    d = bounds.shape[1]
    lo, hi = bounds[0], bounds[1]
    F = len(acq_functions)
    nb = min(num_restarts, raw_samples)
    eng = torch.quasirandom.SobolEngine(d, scramble=True, seed=seed)
    x0 = []
    with torch.no_grad():
        for fn in acq_functions:          # botorch draws fresh raw samples per call; so does every function here
            X = lo + (hi - lo) * eng.draw(raw_samples).to(dev).double()
            vals = fn(X[:, None, :]).double()
            x0.append(X[torch.topk(vals, nb).indices])
    x0 = torch.stack(x0).detach().cpu().numpy().reshape(-1)            # (F, nb, d)

    def joint(Xc):
        return torch.cat([fn(Xc[f * nb:(f + 1) * nb]).double() for f, fn in enumerate(acq_functions)])

    if use_cuda_graph is None:
        use_cuda_graph = dev.type == "cuda"
    info = {"graph": bool(use_cuda_graph), "evaluations": 0}
    if use_cuda_graph:
        graphed = GraphedValueAndGrad(joint, F * nb, d, dev)

        def fun(flat):
            info["evaluations"] += 1
            loss, g, _ = graphed(flat)
            return loss, g
    else:
        def fun(flat):
            info["evaluations"] += 1
            Xc = torch.tensor(flat.reshape(F * nb, 1, d), dtype=torch.double, device=dev, requires_grad=True)
            loss = -joint(Xc).sum()
            g, = torch.autograd.grad(loss, Xc)
            return float(loss.detach()), g.detach().cpu().numpy().reshape(-1).astype(np.float64)

    bnds = list(zip(lo.cpu().numpy().tolist(), hi.cpu().numpy().tolist())) * (F * nb)
    res = minimize(fun, x0, jac=True, method="L-BFGS-B", bounds=bnds,
                   options={"maxiter": options.get("maxiter", 200)})
    Xf = torch.tensor(res.x.reshape(F * nb, 1, d), dtype=torch.double, device=dev)
    with torch.no_grad():
        vf = joint(Xf)
    out = []
    for f in range(F):
        best = int(torch.argmax(vf[f * nb:(f + 1) * nb])) + f * nb
        out.append((Xf[best], vf[best]))
    info["iterations"] = int(res.nit)
    return (out, info) if return_info else out


def optimize_acqf(acq_function, bounds, q=1, num_restarts=5, raw_samples=200, options=None, seed=None,
                  use_cuda_graph=None):
    return optimize_acqf_multi([acq_function], bounds, q=q, num_restarts=num_restarts, raw_samples=raw_samples,
                               options=options, seed=seed, use_cuda_graph=use_cuda_graph)[0]
