"""Minimal stand-in for ``botorch.optim.optimize_acqf`` as called at
``mobocmf/acquisition_functions/JESMOC_MFDGP.py:142-143,159-160`` (q=1): quasi-random raw samples evaluated in one
batch under no_grad, the best ``num_restarts`` refined jointly by L-BFGS-B (scipy) with autograd gradients.  All
restarts are evaluated in ONE acquisition call per L-BFGS iteration so the GPU sees (num_restarts, 1, d) batches."""
import numpy as np
import torch
from scipy.optimize import minimize


def optimize_acqf(acq_function, bounds, q=1, num_restarts=5, raw_samples=200, options=None, seed=None):
    assert q == 1
    options = options or {}
    bounds = torch.as_tensor(bounds, dtype=torch.double)
    dev = bounds.device
    d = bounds.shape[1]
    lo, hi = bounds[0], bounds[1]
    eng = torch.quasirandom.SobolEngine(d, scramble=True, seed=seed)
    X = lo + (hi - lo) * eng.draw(raw_samples).to(dev).double()
    with torch.no_grad():
        vals = acq_function(X[:, None, :]).double()
    idx = torch.topk(vals, min(num_restarts, raw_samples)).indices
    x0 = X[idx].detach().cpu().numpy().reshape(-1)
    nb = len(idx)

    def fun(flat):
        Xc = torch.tensor(flat.reshape(nb, 1, d), dtype=torch.double, device=dev, requires_grad=True)
        v = acq_function(Xc).double()
        loss = -v.sum()
        g, = torch.autograd.grad(loss, Xc)
        return float(loss.detach()), g.detach().cpu().numpy().reshape(-1).astype(np.float64)

    bnds = list(zip(lo.cpu().numpy().tolist(), hi.cpu().numpy().tolist())) * nb
    res = minimize(fun, x0, jac=True, method="L-BFGS-B", bounds=bnds,
                   options={"maxiter": options.get("maxiter", 200)})
    Xf = torch.tensor(res.x.reshape(nb, 1, d), dtype=torch.double, device=dev)
    with torch.no_grad():
        vf = acq_function(Xf).double()
    best = int(torch.argmax(vf))
    return Xf[best], vf[best]
