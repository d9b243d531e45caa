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


def plan_ragged_tiles(steps_desc, state_budget: int = 400_000_000, granule: int = 148 * 128, max_tracks: int = 8 * 148 * 128):
    """Cut a fleet, ordered by DECREASING number of filter steps, into tiles ``[(lo, hi), ...]``.

    A tile's buffers are rectangular, ``(longest track + 1) x tracks`` states, so tiles are sized by
    a budget of stored states (each costs 264-376 bytes across the result arrays and the smoother
    tape): few, long tracks in the first tiles, many short ones in the last.  Tile widths are
    multiples of ``granule`` (one block of 128 tracks per SM of a B200) up to ``max_tracks``; the
    last tile takes what is left.  With the fleet sorted, a tile's tracks differ little in length
    (1 M tracks of 100-5 000 fixes: < 3 % inside any tile), so the rectangle wastes almost nothing
    and warps retire together.  ``shard_tiles_round_robin`` deals the tiles to ranks."""
    steps = [int(v) for v in steps_desc]
    if any(a < b for a, b in zip(steps[:-1], steps[1:])):
        raise ValueError("plan_ragged_tiles needs the tracks in order of decreasing length")
    tiles, lo, n = [], 0, len(steps)
    while lo < n:
        width = state_budget // (steps[lo] + 1)
        width = max(granule, min(max_tracks, (width // granule) * granule))
        hi = min(n, lo + width)
        tiles.append((lo, hi))
        lo = hi
    return tiles


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
