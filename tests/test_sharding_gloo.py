"""CPU, world_size = 2 over gloo: the multi-GPU host logic (track sharding with no data-path
collective, and the optional end-of-job summary reduction) without any GPU."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ship_track_estimators_b200.sharding import reduce_summary, shard_range, shard_tiles_round_robin


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 16, 1000, 16 * 1024 * 1024):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)
    tiles = [shard_tiles_round_robin(11, r, 4) for r in range(4)]
    assert sorted(sum(tiles, [])) == list(range(11))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard_range(1001, rank, world)
        # each rank "processes" its own tracks: the summary counts what it owned
        local = {"tracks": float(hi - lo), "track_steps": float((hi - lo) * 1024), "flagged": float(rank), "updates": 3.0,
                 "checksum": float(sum(range(lo, hi)))}
        total = reduce_summary(local)
        if rank == 0:
            torch.save(total, out)
    finally:
        dist.destroy_process_group()


def test_summary_reduction_world_size_2(tmp_path):
    out = str(tmp_path / "summary.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    total = torch.load(out)
    assert total["tracks"] == 1001.0 and total["track_steps"] == 1001.0 * 1024
    assert total["flagged"] == 1.0 and total["updates"] == 6.0
    assert total["checksum"] == float(sum(range(1001)))


def test_reduce_summary_is_identity_without_process_group():
    assert reduce_summary({"a": 1.5, "b": 2.0}) == {"a": 1.5, "b": 2.0}
