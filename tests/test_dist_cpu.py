"""CPU, world_size 2 over gloo: batch sharding by independent tiles and the optional final gather."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bfsr_b200.dist import gather_tiles, shard_range


def test_shard_range_partitions_exactly():
    for n in (0, 1, 5, 32, 33, 256):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _worker(rank, world, port, n):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = torch.arange(n * 3 * 4 * 4, dtype=torch.float32).view(n, 3, 4, 4)
        lo, hi = shard_range(n, world, rank)
        local = full[lo:hi] * 2.0            # stand-in for the per-rank SR result
        got = gather_tiles(local, n)
        assert torch.equal(got, full * 2.0)
        got0 = gather_tiles(local, n, dst=0)
        assert (got0 is None) == (rank != 0)
        if rank == 0:
            assert torch.equal(got0, full * 2.0)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [5, 8])
def test_gather_tiles_world2_gloo(n):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, n), nprocs=2, join=True)
