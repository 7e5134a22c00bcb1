"""Multi-GPU plumbing: the hot path shards by independent LR tiles, one process per GPU, no data-path collective.

The reference's only parallelism is `nn.DataParallel` (SRFlow_model.py:53), which re-broadcasts the weights on every
forward; here every rank holds its own packed weights and processes a contiguous slice of the batch.  The optional final
gather of the SR tiles is the single collective (NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n: int, world: int, rank: int):
    """Contiguous, balanced slice [lo, hi) of n independent tiles for `rank` (first n % world ranks get one extra)."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError(f"bad rank {rank} / world {world}")
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def bucket_by_scale(scales, world: int = 1, rank: int = 0):
    """Mixed-scale batches (BASELINE config 5: LINF-LP at scales {2,3,4,6,8}): images of one scale form a bucket (one query grid
    per bucket) and EVERY bucket is split into `world` contiguous shares, so each rank gets the same mix and therefore (up to the
    rotated remainders) the same cost (work grows with the square of the scale).  Returns {scale: [image indices of this rank]} in
    ascending scale order; empty shares are omitted."""
    buckets = {}
    for i, s in enumerate(scales):
        buckets.setdefault(s, []).append(i)
    out = {}
    offset = 0          # ranks that already received an extra image from the smaller-scale buckets
    for s in sorted(buckets):
        idx = buckets[s]
        # rotate the ranks that take the `len % world` extras from bucket to bucket: with 5 buckets of 9 images on 8 ranks no
        # rank gets more than one extra instead of rank 0 getting all five (its step time would set the job's)
        lo, hi = shard_range(len(idx), world, (rank - offset) % world)
        offset = (offset + len(idx) % world) % world
        if hi > lo:
            out[s] = idx[lo:hi]
    return out


def gather_tiles(local: torch.Tensor, n_total: int, group=None, dst=None):
    """Reassemble the per-rank slices (dim 0) of a sharded batch.  dst=None -> every rank gets the full batch
    (all_gather); dst=r -> only rank r does (gather); other ranks get None.  Ragged slices are padded to the largest."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [shard_range(n_total, world, r) for r in range(world)]
    mx = max(hi - lo for lo, hi in sizes)
    buf = local
    if local.shape[0] < mx:
        pad = torch.zeros((mx - local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        buf = torch.cat([local, pad], 0)
    buf = buf.contiguous()
    if dst is None:
        outs = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(outs, buf, group=group)
    else:
        outs = [torch.empty_like(buf) for _ in range(world)] if rank == dst else None
        dist.gather(buf, outs, dst=dst, group=group)
        if rank != dst:
            return None
    return torch.cat([o[:hi - lo] for o, (lo, hi) in zip(outs, sizes)], 0)
