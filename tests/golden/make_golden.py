"""Generates the committed golden vectors tests/golden/*.npz from the CPU oracle (run from the repo root:
``python tests/golden/make_golden.py``).  The reference itself cannot run here (GPyTorch / BoTorch are not installable,
SURVEY.md section 8c), so these vectors pin the ORACLE, not the reference: they catch any drift of either side
(oracle edits, torch upgrades, kernel changes)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import mfdgp_oracle as O                      # noqa: E402
from tests.helpers import random_state, clone_state, param_keys   # noqa: E402

CASES = {
    # name: (M, d, L, B, S, lengthscale, seed)
    "l2_single_sample": (12, 2, 2, 9, 1, 0.35, 11),
    "l3_three_samples": (20, 3, 3, 11, 3, 0.30, 12),
    "l2_forrester_sized": (16, 1, 2, 16, 1, 0.10, 13),
}


def build(name):
    M, d, L, B, S, ls, seed = CASES[name]
    sd, noise_upper = random_state(M, d, L, seed=seed, ls=ls)
    g = torch.Generator().manual_seed(seed + 100)
    x = torch.rand(B, d, generator=g, dtype=torch.float64)
    y = torch.randn(B, 1, generator=g, dtype=torch.float64)
    fid = torch.randint(0, L, (B, 1), generator=g).double()
    eps = [None] + [torch.randn(B * S, generator=g).double() for _ in range(1, L)]
    samples = [torch.randn(5, 1, generator=g) for _ in range(L)]          # float32 like layers/...py:161
    X = torch.rand(7, 1, d, generator=g, dtype=torch.float64)
    sd_c = clone_state(sd)
    for k in sd_c:
        if "chol_variational_covar" in k:
            sd_c[k] = sd_c[k] * 0.6
        if "variational_mean" in k:
            sd_c[k] = sd_c[k] + 0.05 * torch.randn(sd_c[k].shape, generator=g, dtype=torch.float64)
    return dict(M=M, d=d, L=L, B=B, S=S, sd=sd, sd_c=sd_c, noise_upper=noise_upper, x=x, y=y, fid=fid, eps=eps,
                samples=samples, X=X, num_data=3 * B)


def evaluate(c):
    """Every number the parity tests compare, from the oracle."""
    L, S = c["L"], c["S"]
    sdo = clone_state(c["sd"], requires_grad=True)
    loss, kl = O.elbo_step_loss_tiled(sdo, L, c["noise_upper"], c["x"], c["y"], c["fid"], c["eps"], c["num_data"], S)
    loss.backward()
    out = {"loss": loss.detach(), "kl_scaled": kl.detach()}
    for k in param_keys(c["sd"]):
        gk = sdo[k].grad
        out["grad/" + k] = torch.tril(gk) if "chol_variational_covar" in k else gk
    mod_u = dict(sd=c["sd"], num_layers=L, noise_upper=c["noise_upper"], samples=c["samples"])
    mod_c = dict(sd=c["sd_c"], num_layers=L, noise_upper=c["noise_upper"], samples=c["samples"])
    for f in range(L):
        mu, var = O.predict_for_acquisition(c["sd"], L, c["noise_upper"], c["samples"], c["X"], f)
        out["acq_mu/%d" % f], out["acq_var/%d" % f] = mu, var
        out["jes/%d" % f] = O.jes_mfdgp(mod_u, mod_c, c["X"], f)
    return out


def main():
    here = os.path.dirname(os.path.abspath(__file__))
    for name in CASES:
        out = evaluate(build(name))
        np.savez(os.path.join(here, name + ".npz"), **{k: v.detach().numpy() for k, v in out.items()})
        print(name, "loss %.12g" % float(out["loss"]), "entries", len(out))


if __name__ == "__main__":
    main()
