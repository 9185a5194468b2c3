"""CPU, world_size 2, gloo: the host logic of the multi-GPU path (mobocmf_b200/util/distributed.py).

"G-way shard + sum == 1-way": each rank computes the oracle ELBO loss of ITS rows, the flat gradient buffers are
summed with one all-reduce and must equal the gradients of the single-process step on the concatenated minibatch.
The oracle stands in for the CUDA step here (there is no GPU in the CPU suite); the same identity is checked on the
B200 by tests/test_gpu_fused_step.py::test_two_shards_sum_to_one_step.
"""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import mfdgp_oracle as O
from tests.helpers import random_state, clone_state, param_keys
from mobocmf_b200.util import distributed as D


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _problem():
    M, d, L, B = 12, 2, 2, 10
    sd, noise_upper = random_state(M, d, L, seed=3, ls=0.4)
    g = torch.Generator().manual_seed(11)
    x = torch.rand(B, d, generator=g, dtype=torch.float64)
    y = torch.randn(B, 1, generator=g, dtype=torch.float64)
    fid = torch.randint(0, L, (B, 1), generator=g).double()
    eps = torch.randn(1, B, generator=g).double()
    return sd, noise_upper, L, x, y, fid, eps, 40


def _loss(sd, noise_upper, L, x, y, fid, eps, num_data):
    loss, _ = O.elbo_step_loss(sd, L, noise_upper, x, y, fid, [None, eps], num_data)
    return loss


def _worker(rank, world_size, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        sd, noise_upper, L, x, y, fid, eps, num_data = _problem()
        keys = param_keys(sd)
        sdr = clone_state(sd, requires_grad=True)
        if rank != 0:                      # ranks start different, the broadcast makes them equal
            with torch.no_grad():
                for k in keys:
                    sdr[k].add_(1.0)

        class Holder(torch.nn.Module):
            def __init__(self):
                super().__init__()
                self.ps = torch.nn.ParameterList([torch.nn.Parameter(sdr[k].detach().clone()) for k in keys])
        h = Holder()
        D.broadcast_parameters(h)
        for k, p in zip(keys, h.ps):
            sdr[k] = p
        flat = D.FlatGrads(list(h.ps))
        assert flat.attached()
        lo, hi = D.shard_bounds(x.shape[0], rank, world_size)
        loss = _loss(sdr, noise_upper, L, x[lo:hi], y[lo:hi], fid[lo:hi], eps[:, lo:hi], num_data)
        grads = torch.autograd.grad(loss, list(h.ps))
        for p, g_ in zip(h.ps, grads):
            p.grad.copy_(g_)               # what the kernels do: write into the flat buffer's slices
        before = flat.flat.clone()
        flat.all_reduce()
        # the bucketed exchange (large blocks first, one tail) is the same sum: the chol blocks lead a second buffer
        big = [p for k, p in zip(keys, h.ps) if "chol_variational_covar" in k][::-1]
        flat2 = D.FlatGrads(list(h.ps), first=big)
        assert flat2.n_first == len(big) == L and flat2.params[0] is big[0] and flat2.attached()
        for p, g_ in zip(h.ps, grads):
            p.grad.copy_(g_)
        assert sum(b.numel() for b in flat2.buckets()) == flat2.flat.numel() == before.numel()
        flat2.all_reduce_overlapped()
        bucketed = {id(p): p.grad.clone() for p in h.ps}
        flat.reattach()
        for p in h.ps:                       # same per-parameter sums as the single all-reduce of the first layout
            o = flat.offsets[[id(q) for q in flat.params].index(id(p))]
            assert torch.equal(bucketed[id(p)].reshape(-1), flat.flat[o:o + p.numel()])
        vals = torch.arange(lo, hi, dtype=torch.float64) * (1.0 if rank == 0 else -1.0)
        full = D.gather_candidate_values(vals, x.shape[0])
        best = D.argmax_over_ranks(vals, lo)
        torch.save({"flat": flat.flat.clone(), "full": full, "best": best, "rank": rank,
                    "p0": h.ps[0].detach().clone()}, os.path.join(out_dir, "r%d.pt" % rank))
    finally:
        dist.destroy_process_group()


def test_two_rank_shard_and_sum_equals_single_step(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0 = torch.load(os.path.join(tmp_path, "r0.pt"))
    r1 = torch.load(os.path.join(tmp_path, "r1.pt"))
    assert torch.equal(r0["flat"], r1["flat"])             # every rank holds the same summed gradient
    assert torch.equal(r0["p0"], r1["p0"])                 # broadcast made the parameters equal
    sd, noise_upper, L, x, y, fid, eps, num_data = _problem()
    keys = param_keys(sd)
    sdr = clone_state(sd, requires_grad=True)
    loss = _loss(sdr, noise_upper, L, x, y, fid, eps, num_data)
    grads = torch.autograd.grad(loss, [sdr[k] for k in keys])
    ref = torch.cat([g.reshape(-1) for g in grads])
    err = float((r0["flat"] - ref).abs().max() / ref.abs().max())
    assert err < 1e-12, err
    # candidate gather / arg-max
    n = x.shape[0]
    lo1, hi1 = D.shard_bounds(n, 1, 2)
    expect = torch.cat([torch.arange(0, lo1, dtype=torch.float64), -torch.arange(lo1, hi1, dtype=torch.float64)])
    assert torch.equal(r0["full"], expect) and torch.equal(r1["full"], expect)
    assert r0["best"] == r1["best"] == (lo1 - 1, float(lo1 - 1))


def test_shard_bounds_tile_exactly():
    for n in (0, 1, 7, 64, 1000003):
        for ws in (1, 2, 3, 8):
            prev = 0
            for r in range(ws):
                lo, hi = D.shard_bounds(n, r, ws)
                assert lo == prev and hi >= lo and hi - lo in (n // ws, n // ws + 1)
                prev = hi
            assert prev == n


def test_flat_grads_alias_and_reattach():
    ps = [torch.nn.Parameter(torch.randn(3, 2, dtype=torch.float64)), torch.nn.Parameter(torch.randn(5, dtype=torch.float64))]
    frozen = torch.nn.Parameter(torch.randn(2, dtype=torch.float64), requires_grad=False)
    flat = D.FlatGrads(ps + [frozen])
    assert flat.flat.numel() == 11 and frozen.grad is None
    ps[1].grad.fill_(2.0)
    assert float(flat.flat[6:].sum()) == 10.0
    ps[0].grad = None
    assert not flat.attached()
    flat.reattach()
    assert flat.attached()
    assert D.world() == (0, 1)
    # leading blocks: the named parameters first, in the given order; buckets = one per leading block + one tail
    f2 = D.FlatGrads(ps + [frozen], first=[ps[1], frozen])
    assert f2.n_first == 1 and f2.params[0] is ps[1] and f2.offsets == [0, 5]
    b = f2.buckets()
    assert [x.numel() for x in b] == [5, 6] and b[0].data_ptr() == ps[1].grad.data_ptr()
    assert f2.all_reduce_overlapped() is f2.flat          # no process group: a no-op
