"""CPU, world_size 2 over gloo: batch sharding by independent tiles and the optional final gather."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bfsr_b200.dist import bucket_by_scale, gather_tiles, shard_range


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


def test_bucket_by_scale_balances_every_scale_over_ranks():
    scales = [[2, 3, 4, 6, 8][i % 5] for i in range(256)]          # BASELINE config 5
    for world in (1, 2, 8):
        seen = {}
        for r in range(world):
            b = bucket_by_scale(scales, world, r)
            assert list(b) == sorted(b)
            for s, idx in b.items():
                assert all(scales[i] == s for i in idx)
                seen.setdefault(s, []).extend(idx)
            sizes = [len(v) for v in b.values()]
            assert max(sizes) - min(sizes) <= 1                      # every rank gets (almost) the same count of every scale
        assert sorted(i for v in seen.values() for i in v) == list(range(256))
    assert bucket_by_scale([4, 4, 2], 4, 3) == {}                     # more ranks than images of any scale
    # remainders rotate over the ranks: 5 buckets of 9 images on 8 ranks -> nobody holds more than one extra image
    scales = [[2, 3, 4, 6, 8][i % 5] for i in range(45)]
    per_rank = [bucket_by_scale(scales, 8, r) for r in range(8)]
    assert sorted(i for b in per_rank for v in b.values() for i in v) == list(range(45))
    assert max(sum(len(v) for v in b.values()) for b in per_rank) == 6
