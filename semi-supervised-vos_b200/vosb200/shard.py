"""Sequence sharding across ranks (one process per GPU).  Sequences are independent, frames inside
a sequence are serial, so the unit of distribution is a whole sequence; there is no collective on
the hot path -- torch.distributed (NCCL on GPUs, gloo in CPU tests) is used only to gather the
per-sequence results at the end."""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import torch
import torch.distributed as dist


def sequence_cost(n_frames: int, n_pixels: int, ref_num: int = 9) -> float:
    """Affinity FLOPs of one sequence up to a constant: sum_t R_t * P^2 with R_t = min(t, ref_num)."""
    ramp = min(max(n_frames - 1, 0), ref_num)
    refs = ramp * (ramp + 1) // 2 + max(n_frames - 1 - ref_num, 0) * ref_num
    return float(refs) * float(n_pixels) ** 2


def assign_lpt(costs: Sequence[float], world_size: int) -> List[List[int]]:
    """Longest-processing-time-first assignment of sequence indices to ranks (deterministic:
    ties broken by index).  Every rank computes the same table; no communication."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    load = [0.0] * world_size
    out: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += costs[i]
    for r in range(world_size):
        out[r].sort()
    return out


def imbalance(costs: Sequence[float], assignment: List[List[int]]) -> float:
    """max rank load / mean rank load (1.0 = perfect): the scaling ceiling of whole-sequence sharding."""
    loads = [sum(costs[i] for i in a) for a in assignment]
    mean = sum(loads) / len(loads)
    return max(loads) / mean if mean > 0 else 1.0


def gather_results(local: Dict[int, torch.Tensor], dst: int = 0) -> Dict[int, torch.Tensor]:
    """Final per-sequence result gather: {sequence index: (T-1,H,W) uint8 masks} from every rank to
    `dst`.  Tensors may live on the GPU (NCCL) or CPU (gloo); shapes differ per sequence, so sizes
    travel first (all_gather_object) and payloads go as one flat uint8 buffer per rank."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return dict(local)
    world, rank = dist.get_world_size(), dist.get_rank()
    keys = sorted(local)
    meta = [(k, tuple(local[k].shape)) for k in keys]
    metas: List[List[Tuple[int, Tuple[int, ...]]]] = [None] * world  # type: ignore[list-item]
    dist.all_gather_object(metas, meta)
    device = local[keys[0]].device if keys else torch.device('cuda' if dist.get_backend() == 'nccl' else 'cpu')
    flat = torch.cat([local[k].reshape(-1) for k in keys]) if keys else torch.empty(0, dtype=torch.uint8, device=device)
    sizes = [sum(int(torch.Size(s).numel()) for _, s in m) for m in metas]
    out: Dict[int, torch.Tensor] = {}
    if rank == dst:
        bufs = [torch.empty(n, dtype=torch.uint8, device=device) for n in sizes]
        bufs[dst] = flat
        reqs = [dist.irecv(bufs[src], src=src) for src in range(world) if src != dst and sizes[src] > 0]
        for q in reqs:
            q.wait()
        for src in range(world):
            off = 0
            for k, shape in metas[src]:
                n = int(torch.Size(shape).numel())
                out[k] = bufs[src][off:off + n].view(shape)
                off += n
    elif flat.numel() > 0:
        dist.send(flat, dst=dst)
    return out
