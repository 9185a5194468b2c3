#!/usr/bin/env python
"""Pins the oracle against the REAL reference: runs fernandezdaniel/MOBOCMF's own GPyTorch / BoTorch path on the golden
cases of tests/golden/make_golden.py and writes ``tests/golden/ref_<case>.npz`` with the same keys as the oracle's
fixtures (loss, kl_scaled, grad/<parameter>, acq_mu/<f>, acq_var/<f>, jes/<f>).

    python tools/make_golden_from_gpytorch.py --reference /path/to/MOBOCMF [--cases l2_single_sample ...]

Needs gpytorch + botorch + linear_operator importable (the reference asks for botorch >= 0.9.0, i.e. gpytorch 1.11 /
linear_operator 0.5.1).  They are NOT in this image (no network, not in /opt/wheelhouse), which is why DESIGN.md calls
the ELBO / JES oracle "parity unpinned"; the day a wheel exists this one command pins it, and
``tests/test_golden.py::test_oracle_reproduces_reference_golden`` (CPU) / the GPU twin pick the files up.

What is run, all of it the reference's unmodified code:
  * ``mobocmf.models.mfdgp.MFDGP`` built on the case's inducing inputs, its parameters then loaded from the case's
    state (GPyTorch state_dict names, the oracle's native format), noise bounds set to the case's;
  * one ``_update_model`` body per MC sample (``model(x)`` under ``num_likelihood_samples(1)``, ``VariationalELBOMF``,
    ``loss.backward()``; mobocmf/util/blackbox_mfdgp_fitter.py:161-168) with ``torch.normal`` replaced for the duration
    of the forward by a function that hands out the case's normals (the reference draws them inside the layer,
    mobocmf/layers/mfdgp_hidden_layer.py:274); S > 1 cases are the mean of S such steps;
  * ``MFDGP.predict_for_acquisition`` and ``_JES_MFDGP.forward`` for every fidelity.
Prints the relative difference oracle-vs-reference per quantity; exits non-zero above --tol.
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def reference_model(mods, c, sd, L):
    """The reference's MFDGP carrying the oracle-format state ``sd`` (same bridge as tests/helpers.model_from_state)."""
    MFDGP = mods["MFDGP"]
    Zx = sd["hidden_layer_0.variational_strategy.inducing_points"]
    M = Zx.shape[0]
    fid = (torch.arange(M) % L).double()[:, None]
    y = torch.linspace(-1.0, 1.0, M, dtype=torch.float64)[:, None]
    model = MFDGP(Zx.clone(), y, fid, L, num_samples_for_acquisition=c["samples"][0].shape[0])
    model.double()
    own = model.state_dict()
    missing = [k for k in sd if k not in own and "inducing_points" not in k]
    if missing:
        raise SystemExit("state keys the reference model does not have: %s" % missing)
    with torch.no_grad():
        for k, v in sd.items():
            if "inducing_points" not in k:
                own[k].copy_(v.reshape(own[k].shape))
        for l in range(L):
            con = getattr(model, "hidden_layer_likelihood_%d" % l).noise_covar.raw_noise_constraint
            con.upper_bound.copy_(torch.as_tensor(c["noise_upper"][l], dtype=torch.float64))
            con.lower_bound.copy_(torch.as_tensor(1e-8, dtype=torch.float64))
            layer = getattr(model, "hidden_layer_%d" % l)
            layer.samples = c["samples"][l].clone()
            layer.num_samples_for_acquisition = c["samples"][l].shape[0]
    return model


class scripted_normals(object):
    """``torch.normal`` hands out the queued tensors (one per layer >= 1, in call order) instead of drawing."""

    def __init__(self, queue):
        self.queue = list(queue)

    def __enter__(self):
        self.orig = torch.normal
        torch.normal = lambda mean, *a, **k: self.queue.pop(0).reshape(mean.shape).to(mean.dtype)
        return self

    def __exit__(self, *exc):
        torch.normal = self.orig
        if not exc[0] and self.queue:
            raise SystemExit("the reference drew fewer normals than the case provides")
        return False


def run_case(mods, name, build, param_keys):
    import gpytorch
    c = build(name)
    L, S, B = c["L"], c["S"], c["B"]
    model = reference_model(mods, c, c["sd"], L)
    elbo = mods["VariationalELBOMF"](model, c["num_data"], L)
    keys = param_keys(c["sd"])
    params = dict(model.named_parameters())
    loss_sum, kl_sum, grads = 0.0, 0.0, {k: torch.zeros_like(params[k]) for k in keys}
    model.train()
    for s in range(S):
        model.zero_grad(set_to_none=True)
        eps_s = [c["eps"][l].reshape(B, S)[:, s] for l in range(1, L)]       # tiled row b * S + s
        with gpytorch.settings.num_likelihood_samples(1), scripted_normals(eps_s):
            out = model(c["x"])
            res = elbo(out, c["y"].T, c["fid"])
        loss = -res[0]
        loss.backward()
        loss_sum, kl_sum = loss_sum + loss.detach(), kl_sum + res[1].detach()
        for k in keys:
            if params[k].grad is not None:
                grads[k] += params[k].grad
    out = {"loss": loss_sum / S, "kl_scaled": kl_sum / S}
    for k in keys:
        g = grads[k] / S
        out["grad/" + k] = torch.tril(g) if "chol_variational_covar" in k else g
    cond = reference_model(mods, c, c["sd_c"], L)
    for f in range(L):
        with torch.no_grad(), gpytorch.settings.num_likelihood_samples(1):
            model.eval()
            mu, var = model.predict_for_acquisition(c["X"], f)
            model.train()
            out["acq_mu/%d" % f], out["acq_var/%d" % f] = mu, var
            out["jes/%d" % f] = mods["_JES_MFDGP"](f, model, cond)(c["X"])
    return {k: v.detach().double().reshape(-1) if v.ndim == 0 else v.detach().double() for k, v in out.items()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference", help="checkout of fernandezdaniel/MOBOCMF")
    ap.add_argument("--cases", nargs="*", default=None)
    ap.add_argument("--tol", type=float, default=1e-8, help="oracle-vs-reference relative difference that fails the run")
    args = ap.parse_args()
    try:
        import gpytorch  # noqa: F401
        import botorch   # noqa: F401
    except ImportError as e:
        raise SystemExit("gpytorch / botorch are not importable here (%s): the oracle stays unpinned for the ELBO / JES "
                         "path; run this where they are installed" % e)
    sys.path.insert(0, args.reference)
    from mobocmf.acquisition_functions.JESMOC_MFDGP import _JES_MFDGP
    from mobocmf.mlls.variational_elbo_mf import VariationalELBOMF
    from mobocmf.models.mfdgp import MFDGP
    from tests.golden.make_golden import CASES, build, evaluate
    from tests.helpers import param_keys, relerr
    mods = {"MFDGP": MFDGP, "VariationalELBOMF": VariationalELBOMF, "_JES_MFDGP": _JES_MFDGP}
    here = os.path.join(ROOT, "tests", "golden")
    worst = 0.0
    for name in (args.cases or sorted(CASES)):
        ref = run_case(mods, name, build, param_keys)
        orc = evaluate(build(name))
        np.savez(os.path.join(here, "ref_" + name + ".npz"), **{k: v.numpy() for k, v in ref.items()})
        for k in sorted(orc):
            e = relerr(orc[k].reshape(-1), ref[k].reshape(-1))
            worst = max(worst, e)
            print("%-22s %-100s oracle vs reference %.2e" % (name, k, e))
    print("worst %.2e (tolerance %.1e)" % (worst, args.tol))
    sys.exit(0 if worst <= args.tol else 1)


if __name__ == "__main__":
    main()
