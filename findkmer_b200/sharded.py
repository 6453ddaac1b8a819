"""Multi-GPU plumbing of the counting path: contiguous shards, a 16-byte left halo,
and ONE exchange step -- sum-reduce of the per-GPU tables, max-reduce of the prefix flags, sum-reduce of the
12 partial scalars -- before rank 0 finalizes.  One process per GPU; torch.distributed is plumbing only.

The reference's only parallelism is one OS process per k (findKmer/k6thru11fullANDupstream.sh:16-24); sharding a
single scan is new, and exact because a window belongs to the shard that owns its LAST byte.
"""
from __future__ import annotations

HALO = 16  # bytes of left context a shard needs (k <= 16)


def shard_cuts(total: int, world: int):
    """world+1 cut points over [0,total): contiguous ranges, interior cuts 16-byte aligned."""
    cuts = [min(total, ((total * r // world) + 15) // 16 * 16) for r in range(world)] + [total]
    cuts[0] = 0
    return cuts


def shard_range(total: int, world: int, rank: int):
    """(begin, end, halo_begin): the shard owns [begin,end) and must be able to read from halo_begin."""
    cuts = shard_cuts(total, world)
    b, e = cuts[rank], cuts[rank + 1]
    return b, e, max(0, b - HALO)


def reduce_accumulators(table, flags, partials, dst: int = 0, group=None, fused=None):
    """The one exchange step.  Tensors may live on CUDA (NCCL) or on the CPU (gloo, tests).
    `fused`: an int32 tensor that holds table and flags back to back (DeviceAccumulators.buf) -- one collective
    instead of two; flag bytes are 0/1 per rank, so their byte-wise sum is exact and "present" == non-zero."""
    import os
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    if fused is not None:
        # FKB_EXCHANGE=allreduce: the same sum delivered to every rank (lets NCCL pick its NVSwitch in-network reduction); measured
        # against the default rooted reduce in profiles/ (the counts are identical either way, only rank `dst` finalizes)
        if os.environ.get("FKB_EXCHANGE", "reduce") == "allreduce":
            dist.all_reduce(fused, op=dist.ReduceOp.SUM, group=group)
        else:
            dist.reduce(fused, dst, op=dist.ReduceOp.SUM, group=group)
    else:
        dist.reduce(table, dst, op=dist.ReduceOp.SUM, group=group)  # uint32 bit patterns in int32: wrap-around sums are exact
        dist.reduce(flags, dst, op=dist.ReduceOp.MAX, group=group)  # presence bytes: OR == max
    dist.reduce(partials, dst, op=dist.ReduceOp.SUM, group=group)


def agree_on_stop(stop_offset_global, group=None):
    """A byte 0xFF outside a header ends the reference's scan (findKmer.cpp:975,:988).  Every rank passes the
    GLOBAL file offset of the first such byte it met (or None); returns the minimum over ranks (or None).
    Ranks whose shard begins after it must discard their accumulators before the reduce."""
    import torch
    import torch.distributed as dist
    big = (1 << 62)
    v = big if stop_offset_global is None else int(stop_offset_global)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
        t = torch.tensor([v], dtype=torch.int64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
        v = int(t[0])
    return None if v >= big else v
