"""Multi-GPU plumbing: image pairs are independent end to end, so a batch is split
contiguously across ranks (one process per GPU, weights replicated) and the ONLY collective is the
final gather of match indices / scores (int32 on the wire, widened to int64 at the boundary)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_pairs: int, rank: int, world: int):
    """Contiguous [lo, hi) of the pairs owned by `rank`; remainder pairs go to the first ranks."""
    base, rem = divmod(n_pairs, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_matches(matches0: torch.Tensor, scores0: torch.Tensor, n_pairs: int, group=None):
    """All-gather per-rank (b_local, N) match indices and scores into (n_pairs, N) on every rank.

    Shards may be uneven by one pair; they are padded to the largest shard for the collective."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return matches0, scores0
    rank = dist.get_rank(group)
    N = matches0.shape[1]
    bmax = (n_pairs + world - 1) // world
    wire_m = torch.full((bmax, N), -1, dtype=torch.int32, device=matches0.device)
    wire_s = torch.zeros((bmax, N), dtype=torch.float32, device=matches0.device)
    wire_m[:matches0.shape[0]] = matches0.to(torch.int32)
    wire_s[:scores0.shape[0]] = scores0
    out_m = torch.empty((world * bmax, N), dtype=torch.int32, device=matches0.device)
    out_s = torch.empty((world * bmax, N), dtype=torch.float32, device=matches0.device)
    dist.all_gather_into_tensor(out_m, wire_m, group=group)
    dist.all_gather_into_tensor(out_s, wire_s, group=group)
    keep = []
    for r in range(world):
        lo, hi = shard_range(n_pairs, r, world)
        keep.append(torch.arange(r * bmax, r * bmax + (hi - lo), device=matches0.device))
    keep = torch.cat(keep)
    return out_m[keep].to(torch.int64), out_s[keep]
