"""Multi-GPU sharding of one FASTQ byte stream.

The stream is cut into contiguous byte shards, one per rank.  A record belongs to the shard that
holds its first byte, so a rank only needs to know where its first own record starts and where
the next rank's first own record starts.  Both follow from the newline-rank algebra
(csrc/tile_math.h): every rank summarises its shard on its GPU (64 bytes), the summaries are
all-gathered, and bsq_shard_prefix (host arithmetic) gives every rank its cut points.  The only
other collective is the all-reduce of the final read / base counts.  No FASTQ bytes cross GPUs
beyond the halo a rank reads past its shard end to finish its last record.

`dist` is torch.distributed (NCCL on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _capi as capi
from .host import shard_prefix


@dataclass
class ShardPlan:
    rank: int
    world: int
    begin: int          # first byte (relative to the rank's shard) of its first own record
    end: int            # one past its last own record, relative to the shard (may exceed the shard: halo)
    first_record: int   # global index of its first own record
    newline_rank: int


def gather_summaries(dist, local: capi.Summary, local_bytes: int, device="cpu", group=None):
    """all_gather of (64-byte summary, shard size) -> lists in rank order."""
    import torch
    world = dist.get_world_size(group)
    mine = np.zeros(18, np.uint32)
    mine[:16] = np.ctypeslib.as_array(local.w)
    mine[16] = local_bytes & 0xFFFFFFFF
    mine[17] = local_bytes >> 32
    t = torch.from_numpy(mine.view(np.int32).copy()).to(device)
    out = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(out, t, group=group)
    sums, sizes = [], []
    for o in out:
        a = o.cpu().numpy().view(np.uint32)
        s = capi.Summary()
        C.memmove(C.byref(s), a[:16].ctypes.data, 64)
        sums.append(s)
        sizes.append(int(a[16]) | (int(a[17]) << 32))
    return sums, sizes


def plan(dist, local: capi.Summary, local_bytes: int, device="cpu", group=None) -> ShardPlan:
    """Cut points of this rank's records inside its shard."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    sums, sizes = gather_summaries(dist, local, local_bytes, device, group)
    st = shard_prefix(sums, sizes)
    begin = int(st[rank].skip_bytes)
    # own records end where the next shard that owns a record starts owning
    end = sizes[rank]
    j = rank + 1
    while j < world:
        end += int(st[j].skip_bytes)
        if st[j].skip_bytes < sizes[j]:
            break
        j += 1
    if begin >= sizes[rank]:   # no record starts in this shard
        begin = end = sizes[rank]
    return ShardPlan(rank, world, begin, end, int(st[rank].first_record), int(st[rank].newline_rank))


def allreduce_counts(dist, reads: int, bases: int, device="cpu", group=None):
    """The final read / base counts: the one data-path-free collective the path needs."""
    import torch
    t = torch.tensor([reads, bases], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return int(t[0]), int(t[1])
