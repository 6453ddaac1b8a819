"""The N > 1 path on CPU: world_size-2 gloo processes exercise shard planning, halo ownership and the one
exchange step (findkmer_b200/sharded.py).  Per-shard tables come from the oracle here (there is no GPU);
on the GPU box tests/test_gpu_parity.py::test_device_ranges_compose runs the same composition with the kernels."""
import os
import socket

import numpy as np
import pytest

from conftest import random_fasta


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, k, stream_bytes, out_path):
    import torch
    import torch.distributed as dist
    from findkmer_b200 import sharded
    from oracle import harness as H
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    stream = np.frombuffer(stream_bytes, dtype=np.uint8)
    b, e, hb = sharded.shard_range(len(stream), world, rank)
    # windows whose LAST byte lies in [b,e), given only the bytes from the halo on: count(hb..e) - count(hb..b)
    upto_e, upto_b = H.oracle_count_stream(stream[hb:e], k), H.oracle_count_stream(stream[hb:b], k)
    table = torch.from_numpy((upto_e.table - upto_b.table).view(np.int32).copy())
    flags = torch.zeros(16, dtype=torch.uint8)
    flags[rank] = 1
    partials = torch.tensor([upto_e.n_kmers - upto_b.n_kmers] + [0] * 11, dtype=torch.int64)
    sharded.reduce_accumulators(table, flags, partials, dst=0)
    stop = sharded.agree_on_stop(12345 if rank == 1 else None)
    if rank == 0:
        np.save(out_path, table.numpy().view(np.uint32))
        assert int(partials[0]) == int(table.numpy().view(np.uint32).sum(dtype=np.uint64))
        assert flags[:world].tolist() == [1] * world
    assert stop == 12345
    dist.destroy_process_group()


@pytest.mark.parametrize("k", [3, 11])
def test_two_rank_gloo_shards_reduce_to_the_whole(harness, tmp_path, k):
    import torch.multiprocessing as mp
    from findkmer_b200.engine import KmerCounter
    stream = KmerCounter.strip(random_fasta(21, 40000))
    out = str(tmp_path / "table.npy")
    mp.spawn(_worker, args=(2, _free_port(), k, bytes(stream), out), nprocs=2, join=True)
    assert np.array_equal(np.load(out), harness.oracle_count_stream(stream, k).table)


def _worker_fused(rank, world, port, mode, out_path):
    import torch
    import torch.distributed as dist
    from findkmer_b200 import sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["FKB_EXCHANGE"] = mode
    dist.init_process_group("gloo", rank=rank, world_size=world)
    fused = (torch.arange(4096 + 16, dtype=torch.int32) * (rank + 1)) % 1000  # table (4096) and flags (16 words) back to back, as DeviceAccumulators.buf
    partials = torch.full((12,), rank + 1, dtype=torch.int64)
    sharded.reduce_accumulators(fused[:4096], fused[4096:], partials, dst=0, fused=fused)
    if rank == 0:
        np.save(out_path, fused.numpy())
        assert partials.tolist() == [sum(range(1, world + 1))] * 12
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["reduce", "allreduce"])
def test_fused_exchange_modes(tmp_path, mode):
    """the one-collective form of the exchange (table + flags in one int32 buffer), rooted reduce and FKB_EXCHANGE=allreduce"""
    import torch.multiprocessing as mp
    out = str(tmp_path / "fused.npy")
    mp.spawn(_worker_fused, args=(2, _free_port(), mode, out), nprocs=2, join=True)
    i = np.arange(4096 + 16, dtype=np.int64)
    assert np.array_equal(np.load(out), ((i % 1000) + (2 * i) % 1000).astype(np.int32))


def test_shard_cuts_cover_and_align():
    from findkmer_b200 import sharded
    for total in (0, 1, 15, 16, 17, 1000, 3_100_000_024):
        for world in (1, 2, 3, 4, 8):
            cuts = sharded.shard_cuts(total, world)
            assert cuts[0] == 0 and cuts[-1] == total and all(a <= b for a, b in zip(cuts, cuts[1:]))
            assert all(c % 16 == 0 or c == total for c in cuts[1:-1])
            for r in range(world):
                b, e, hb = sharded.shard_range(total, world, r)
                assert hb == max(0, b - 16)
