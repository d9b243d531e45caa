"""Multi-GPU partitioning of a track fleet: tracks are independent, so they shard with NO
collective on the data path.  One process per GPU (``torchrun``); each rank
filters and smooths its own contiguous range of (length-sorted) tracks and keeps its states on
its own device.  The only exchange is an optional, latency-bound reduction of a handful of
summary numbers at the very end (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch


def shard_range(n_items: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous ``[lo, hi)`` slice of ``n_items`` owned by ``rank``; sizes differ by at most one."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of {world_size}")
    base, extra = divmod(int(n_items), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_tiles_round_robin(n_tiles: int, rank: int, world_size: int):
    """Tile indices for ``rank`` when tiles are ordered by decreasing track length (ragged fleets):
    dealing them round-robin balances the work without any communication."""
    return list(range(rank, int(n_tiles), int(world_size)))


def local_summary(res, track_steps: int) -> Dict[str, float]:
    """Per-rank summary of a finished tile: counts and a cheap checksum of the final states."""
    last = res.mean_s if res.mean_s is not None else res.mean_f
    return {
        "track_steps": float(track_steps),
        "tracks": float(res.status.numel()),
        "flagged": float((res.status != 0).sum().item()),
        "updates": float(res.n_updates.sum().item()),
        "checksum": float(torch.nan_to_num(last[0]).sum().item()),
    }


def reduce_summary(summary: Dict[str, float], device=None) -> Dict[str, float]:
    """Sum every entry over all ranks (no-op without an initialised process group)."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return dict(summary)
    keys = sorted(summary)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    buf = torch.tensor([summary[k] for k in keys], dtype=torch.float64, device=device)
    dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    return {k: float(v) for k, v in zip(keys, buf.tolist())}
