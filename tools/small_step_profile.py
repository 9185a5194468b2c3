"""Per-kernel CUDA-event times of one fused ELBO step at the reference's Forrester size (config C2: d=1, 2 fidelities,
N=M=B=16), where a step is pure launch latency.  Run on a GPU box: python tools/small_step_profile.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mobocmf_b200 import _lib
from mobocmf_b200.fused import FusedELBOStep, Adam
from mobocmf_b200.mlls.variational_elbo_mf import VariationalELBOMF
from mobocmf_b200.models.mfdgp import MFDGP
from tests.helpers import forrester_data
dev = torch.device("cuda:0")
x, ys, fid = forrester_data()
torch.manual_seed(0)
model = MFDGP(x, ys["obj1"], fid, 2); model.double().to(dev)
elbo = VariationalELBOMF(model, 16, 2)
step = FusedELBOStep(model, elbo)
opt = Adam(model.parameters(), lr=1e-3)
perm = torch.randperm(16)
xb, yb, fb = x[perm].to(dev), ys["obj1"][perm].to(dev), fid[perm].to(dev)
for _ in range(5):
    step(xb, yb, fb); opt.step()
torch.cuda.synchronize()
_lib.load().mobo_step_side_stream(0)
_lib.profile_enable(True)
for _ in range(20):
    step(xb, yb, fb); opt.step()
torch.cuda.synchronize()
tot = {}
order = []
for name, t in _lib.profile_collect():
    if name not in tot: order.append(name)
    tot.setdefault(name, []).append(t)
_lib.profile_enable(False)
s = 0
for k in order:
    v = tot[k]; s += sum(v) / 20
    print("%-28s n/step %4.1f  us/launch %6.1f  us/step %6.1f" % (k, len(v) / 20, 1e3 * sum(v) / len(v), 1e3 * sum(v) / 20))
print("sum us/step", 1e3 * s)
