"""Host-side logic of the multi-GPU path: one process per GPU, rows sharded, ONE all-reduce of the
small parameter gradients.  The reference has no distributed code at all; this is the B200 scaling
axis of its ``_update_model`` loop (``mobocmf/util/blackbox_mfdgp_fitter.py:156-173``) and of ``coupled_acq``
(``mobocmf/acquisition_functions/JESMOC_MFDGP.py:125-135``).

Semantics.  Rank r runs the ELBO step on its own minibatch of B_r rows: loss_r = -(data_r - KL * B_r / N).  Summing
the gradients over ranks gives the gradient of -(sum_r data_r - KL * (sum_r B_r) / N), i.e. exactly the
single-process step on the concatenated minibatch, so "G-way shard + sum == 1-way" holds to summation order.
The M x M operator chain and the KL are recomputed on every rank (deterministic, identical); Adam then runs
replicated on identical gradients, so parameters never need a broadcast after the initial one.

Nothing here touches CUDA directly, so it is tested on CPU with the gloo backend (tests/test_distributed_gloo.py).
"""
import torch
import torch.distributed as dist


def world():
    """(rank, world_size) of the default process group, (0, 1) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(n, rank, world_size):
    """Contiguous, balanced partition of n units (rows / candidates): rank r owns [lo, hi).  The first n % G ranks
    get one extra unit; the shards tile [0, n) exactly."""
    base, extra = divmod(n, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class FlatGrads(object):
    """One contiguous gradient buffer whose slices are the ``.grad`` of the given parameters, so that the kernels
    write every gradient straight into it and the cross-rank exchange is ONE all-reduce without packing copies."""

    def __init__(self, params, first=()):
        """``first``: parameters whose gradients are laid out at the front of the buffer, in this order (the large
        M x M blocks that ``all_reduce_overlapped`` sends early); the rest follow in the order of ``params``."""
        ps = [p for p in params if p.requires_grad]
        head = [p for p in first if p.requires_grad and any(p is q for q in ps)]
        self.params = head + [p for p in ps if not any(p is q for q in head)]
        self.n_first = len(head)
        if not self.params:
            raise ValueError("no trainable parameters")
        p0 = self.params[0]
        self.offsets = []
        off = 0
        for p in self.params:
            self.offsets.append(off)
            off += p.numel()
        self.flat = torch.zeros(off, dtype=p0.dtype, device=p0.device)
        for p, o in zip(self.params, self.offsets):
            p.grad = self.flat[o:o + p.numel()].view(p.shape)

    def attached(self):
        """True while every parameter's .grad still aliases the flat buffer (``zero_grad(set_to_none=True)`` or an
        optimiser that replaces .grad breaks the aliasing)."""
        return all(p.grad is not None and p.grad.data_ptr() == self.flat.data_ptr() + o * self.flat.element_size()
                   for p, o in zip(self.params, self.offsets))

    def reattach(self):
        for p, o in zip(self.params, self.offsets):
            p.grad = self.flat[o:o + p.numel()].view(p.shape)

    def zero_(self):
        self.flat.zero_()

    def all_reduce(self, group=None):
        """Sum the gradients over the ranks (NCCL over NVLink on the GPU box, gloo in the CPU tests)."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
        return self.flat

    def buckets(self):
        """[(view of the flat buffer)] : one per leading block (``first``), then ONE for everything else."""
        out = []
        for p, o in zip(self.params[:self.n_first], self.offsets[:self.n_first]):
            out.append(self.flat[o:o + p.numel()])
        tail = self.offsets[self.n_first] if self.n_first < len(self.params) else self.flat.numel()
        out.append(self.flat[tail:])
        return out

    def all_reduce_overlapped(self, ready=None, comm_stream=None, group=None):
        """The same sum as ``all_reduce`` in ``n_first + 1`` pieces: leading block k is sent as soon as ``ready(k,
        stream)`` has made the communication stream wait for it (FusedELBOStep.wait_bucket: the event the fused step
        records when that layer's d L_q is final), i.e. behind the rest of the step; only the small tail waits for the
        end of the step.  ``ready`` / ``comm_stream`` None (CPU, gloo): the pieces are sent in order on the spot."""
        if not (dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1):
            return self.flat
        bs = self.buckets()
        if comm_stream is None:
            for b in bs:
                if b.numel():
                    dist.all_reduce(b, op=dist.ReduceOp.SUM, group=group)
            return self.flat
        main = torch.cuda.current_stream(self.flat.device)
        works = []
        for k, b in enumerate(bs[:-1]):
            ready(k, comm_stream)
            with torch.cuda.stream(comm_stream):
                works.append(dist.all_reduce(b, op=dist.ReduceOp.SUM, group=group, async_op=True))
        if bs[-1].numel():
            dist.all_reduce(bs[-1], op=dist.ReduceOp.SUM, group=group)     # on the main stream: after the step
        for w in works:
            w.wait()                                                        # main stream waits for the early pieces
        return self.flat


def broadcast_parameters(module, src=0, group=None):
    """Make every rank start from rank ``src``'s parameters, buffers AND fixed eval-mode normals.  The latter are
    drawn at construction (``samples``, layers/mfdgp_hidden_layer.py:161) and are a plain attribute, not a buffer
    (as in the reference, so that state_dicts stay compatible): without this, ranks that were seeded differently
    would evaluate their acquisition shards with different normals and ``gather_candidate_values`` /
    ``argmax_over_ranks`` would compare inconsistent values."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src, group=group)
    for sub in module.modules():
        smp = getattr(sub, "samples", None)
        if torch.is_tensor(smp):
            if dist.get_backend(group) == "nccl" and not smp.is_cuda:
                dev = next((p.device for p in module.parameters() if p.is_cuda), None)
                buf = smp.to(dev)
                dist.broadcast(buf, src=src, group=group)
                smp.copy_(buf.cpu())
            else:
                dist.broadcast(smp, src=src, group=group)
            cache = getattr(sub, "_dev_cache", None)
            if isinstance(cache, dict):          # device copies of the old samples
                for k in [k for k in cache if isinstance(k, tuple) and k and k[0] == "samples"]:
                    del cache[k]


def gather_candidate_values(local_values, n_total, group=None):
    """Acquisition sweep: every rank evaluated its ``shard_bounds`` slice of the n_total candidates; returns the full
    (n_total,) vector on every rank (padded all-gather, no data-path collective before this point)."""
    rank, ws = world()
    if ws == 1:
        return local_values
    sizes = [shard_bounds(n_total, r, ws)[1] - shard_bounds(n_total, r, ws)[0] for r in range(ws)]
    mx = max(sizes)
    pad = torch.zeros(mx, dtype=local_values.dtype, device=local_values.device)
    pad[:local_values.numel()] = local_values
    out = [torch.empty_like(pad) for _ in range(ws)]
    dist.all_gather(out, pad, group=group)
    return torch.cat([o[:s] for o, s in zip(out, sizes)])


def argmax_over_ranks(local_values, lo, group=None):
    """(global index, value) of the best candidate over all shards; ties resolved to the lowest global index so the
    result does not depend on the number of ranks."""
    rank, ws = world()
    v, i = torch.max(local_values, 0)
    best = torch.stack([v.double(), (i + lo).double()])
    if ws == 1:
        return int(best[1]), float(best[0])
    allb = [torch.empty_like(best) for _ in range(ws)]
    dist.all_gather(allb, best, group=group)
    allb = torch.stack(allb).cpu()
    vmax = allb[:, 0].max()
    idx = allb[allb[:, 0] == vmax, 1].min()
    return int(idx), float(vmax)
