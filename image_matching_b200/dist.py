"""Multi-GPU plumbing: image pairs are independent end to end, so a batch is split
contiguously across ranks (one process per GPU, weights replicated) and the ONLY collective is the
final gather of match indices / scores.

Wire format: ONE int32 buffer per rank, (b_wire, 2, N) -- per pair a row of match indices (int64 -> int32)
and a row of the matching scores' bit patterns -- so a single ``all_gather_into_tensor`` moves both;
indices are widened back to int64 at the boundary.  On CUDA tensors the pack / unpack steps are one
kernel each (``b200m_pack_match_wire`` / ``b200m_unpack_match_wire``: no index glue, no intermediate
casts); CPU tensors (the world-size-2 gloo tests of this host logic) use the equivalent torch ops."""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist


def shard_range(n_pairs: int, rank: int, world: int):
    """Contiguous [lo, hi) of the pairs owned by `rank`; remainder pairs go to the first ranks."""
    base, rem = divmod(n_pairs, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _pack(matches0, scores0, b_wire, handle):
    b, N = matches0.shape
    wire = torch.empty((b_wire, 2, N), dtype=torch.int32, device=matches0.device)
    if matches0.is_cuda:
        from . import lib as _lib
        if matches0.dtype != torch.int64 or scores0.dtype != torch.float32 or matches0.stride(1) != 1 \
                or scores0.stride(1) != 1 or matches0.stride(0) != scores0.stride(0):
            matches0, scores0 = matches0.to(torch.int64).contiguous(), scores0.float().contiguous()
        st = C.c_void_p(torch.cuda.current_stream(matches0.device).cuda_stream)
        _lib.check(_lib.load().b200m_pack_match_wire(handle, C.c_void_p(matches0.data_ptr()),
                                                     C.c_void_p(scores0.data_ptr()), b, b_wire, N,
                                                     matches0.stride(0) if b > 1 else N, C.c_void_p(wire.data_ptr()), st),
                   "b200m_pack_match_wire")
        return wire
    wire[:, 0] = -1
    wire[:, 1] = 0
    wire[:b, 0] = matches0.to(torch.int32)
    wire[:b, 1] = scores0.float().contiguous().view(torch.int32)
    return wire


def _unpack(gathered, world, b_wire, n_pairs, handle):
    N = gathered.shape[-1]
    if gathered.is_cuda:
        from . import lib as _lib
        m = torch.empty((n_pairs, N), dtype=torch.int64, device=gathered.device)
        s = torch.empty((n_pairs, N), dtype=torch.float32, device=gathered.device)
        st = C.c_void_p(torch.cuda.current_stream(gathered.device).cuda_stream)
        _lib.check(_lib.load().b200m_unpack_match_wire(handle, C.c_void_p(gathered.data_ptr()), world, b_wire, n_pairs,
                                                       N, C.c_void_p(m.data_ptr()), C.c_void_p(s.data_ptr()), st),
                   "b200m_unpack_match_wire")
        return m, s
    g = gathered.view(world, b_wire, 2, N)
    parts = [g[r, :shard_range(n_pairs, r, world)[1] - shard_range(n_pairs, r, world)[0]] for r in range(world)]
    g = torch.cat(parts)
    return g[:, 0].to(torch.int64), g[:, 1].contiguous().view(torch.float32)


def gather_matches(matches0: torch.Tensor, scores0: torch.Tensor, n_pairs: int, group=None, handle=None,
                   dst: int | None = None):
    """Gather per-rank (b_local, N) match indices and scores into (n_pairs, N) with ONE collective.

    Shards may be uneven by one pair (padded to the largest shard on the wire).  ``handle``: the b200m handle of this
    rank's model (``Matching._engine.handle``), required for CUDA tensors.  ``dst`` = None: every rank gets the result
    (all-gather); ``dst`` = r: only rank r unpacks and returns it, the others return ``(None, None)`` -- the collective
    is the same (NCCL's all-gather is the cheapest way to get 8 x ~1 MB onto one GPU over NVSwitch), but the other
    ranks skip the unpack kernel."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return matches0, scores0
    if matches0.is_cuda and handle is None:
        raise ValueError("gather_matches on CUDA tensors needs the model's b200m handle")
    N = matches0.shape[1]
    b_wire = (n_pairs + world - 1) // world
    wire = _pack(matches0, scores0, b_wire, handle)
    out = torch.empty((world * b_wire, 2, N), dtype=torch.int32, device=matches0.device)
    dist.all_gather_into_tensor(out, wire, group=group)
    if dst is not None and dist.get_rank(group) != dst:
        return None, None
    return _unpack(out, world, b_wire, n_pairs, handle)
