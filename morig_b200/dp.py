"""Data-parallel plumbing: whole meshes per rank, no collective in the forward (meshes only interact
through per-graph pooling, models/rignet.py:63, and eval-mode BatchNorm constants).  NCCL (or gloo on CPU
test rigs) is used for rendezvous, timing reductions and, when asked, for gathering outputs."""
from __future__ import annotations

from typing import List, Sequence

import torch
import torch.distributed as dist


def partition(costs: Sequence[int], world: int) -> List[List[int]]:
    """Greedy longest-processing-time assignment of mesh indices to ranks by cost (E_tpl + E_geo).
    Deterministic: ties go to the lower rank; indices inside a rank stay in input order."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    load = [0] * world
    bins: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        bins[r].append(i)
        load[r] += costs[i]
    return [sorted(b) for b in bins]


def mesh_cost(mesh: dict) -> int:
    return int(mesh["tpl_edge_index"].shape[1] + mesh["geo_edge_index"].shape[1])


def max_over_ranks(value: float, device) -> float:
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_rows(local: torch.Tensor, rows_per_rank: Sequence[int]) -> torch.Tensor:
    """all_gather of per-rank output rows (ragged) -> rank-ordered concatenation on every rank."""
    world = dist.get_world_size()
    m = max(rows_per_rank)
    pad = torch.zeros((m,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    return torch.cat([b[:n] for b, n in zip(bufs, rows_per_rank)], dim=0)


class GradAllReduce:
    """Data-parallel training: ONE flat fp32 gradient buffer for the whole network (7.88 M floats =
    31.5 MB for jointnet / masknet) all-reduced over NCCL / NVLink once per step -- the only collective of the path.

    `param.grad` of every parameter is a view into the flat buffer, so autograd accumulates straight into it and no
    gather / scatter copies exist.  The buffer is split into buckets in reverse registration order (the order the
    backward finishes them: task head first, then the aggregator, the shared motion encoder last); a bucket's
    all-reduce starts on a side stream as soon as its last gradient has been accumulated, overlapping the rest of the
    backward.  `finish()` waits for the reductions and averages.

        dp = GradAllReduce(model)                 # after dist.init_process_group
        for batch in loader:
            dp.zero_grad()
            loss(model(batch, flow)).backward()
            dp.finish()                           # gradients are now the mean over ranks
            optimizer.step()

    BatchNorm batch statistics stay per replica (as torch's DistributedDataParallel without SyncBatchNorm; the
    reference itself is single-GPU, training/train_rig.py:16)."""

    def __init__(self, model: torch.nn.Module, bucket_mb: float = 12.0, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        params = [p for p in model.parameters() if p.requires_grad]
        if not params:
            raise ValueError("no trainable parameters")
        dev, dtype = params[0].device, params[0].dtype
        order = list(reversed(params))                         # backward completion order, roughly
        total = sum(p.numel() for p in order)
        self.flat = torch.zeros(total, dtype=dtype, device=dev)
        self.params = order
        self.buckets = []                                      # (start, end, number of parameters)
        off = start = count = 0
        limit = int(bucket_mb * (1 << 20) / self.flat.element_size())
        self._bucket_of = {}
        for p in order:
            self._bucket_of[id(p)] = len(self.buckets)
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()
            count += 1
            if off - start >= limit:
                self.buckets.append((start, off, count))
                start, count = off, 0
        if count:
            self.buckets.append((start, off, count))
        self._pending = [0] * len(self.buckets)
        self._work = []
        self._stream = torch.cuda.Stream(dev) if dev.type == "cuda" else None
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in order]
        self._arm()

    def _arm(self):
        self._pending = [b[2] for b in self.buckets]
        self._work = []

    def zero_grad(self):
        self.flat.zero_()
        off = 0
        for p in self.params:                                  # re-attach (optimizer.zero_grad(set_to_none=True) drops them)
            if p.grad is None or p.grad.data_ptr() != self.flat.data_ptr() + off * self.flat.element_size():
                p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        self._arm()

    def _on_grad(self, p):
        b = self._bucket_of[id(p)]
        self._pending[b] -= 1
        if self._pending[b] == 0 and self.world > 1:
            self._launch(b)

    def _launch(self, b):
        start, end, _ = self.buckets[b]
        chunk = self.flat[start:end]
        if self._stream is not None:
            self._stream.wait_stream(torch.cuda.current_stream(self.flat.device))
            with torch.cuda.stream(self._stream):
                self._work.append(dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        else:
            self._work.append(dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def finish(self):
        """wait for every bucket and turn the sums into means; returns the bytes all-reduced this step"""
        if self.world > 1:
            for b, left in enumerate(self._pending):           # parameters that took no part in this backward
                if left > 0:
                    self._launch(b)
            for w in self._work:
                w.wait()
            if self._stream is not None:
                torch.cuda.current_stream(self.flat.device).wait_stream(self._stream)
            self.flat.div_(self.world)
        nbytes = self.flat.numel() * self.flat.element_size() if self.world > 1 else 0
        self._arm()
        return nbytes

    def close(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []
