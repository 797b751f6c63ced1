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
