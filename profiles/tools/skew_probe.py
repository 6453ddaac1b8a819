#!/usr/bin/env python
"""Throughput of the count path on a base-composition-skewed stream (human-like: A = T = 0.295, C = G = 0.205, i.i.d.),
next to the uniform one of the same length -- real genomes load the 1024 buckets unevenly (an A/T-rich 5-base core is ~6x
as frequent as a C/G-rich one), which the uniform synthetic workload of the bench does not show.
    python profiles/tools/skew_probe.py [n_bases] [k,...]"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import numpy as np
import torch
from findkmer_b200.engine import KmerCounter

n = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000_000
ks = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [11, 8]
c = KmerCounter(0)
lib, ctx = c._lib, c._ctx
st = torch.cuda.current_stream()
letters = torch.tensor([65, 67, 71, 84], dtype=torch.uint8, device="cuda")
g = torch.Generator(device="cuda").manual_seed(7)
for name, probs in (("uniform", [0.25, 0.25, 0.25, 0.25]), ("human-like composition", [0.295, 0.205, 0.205, 0.295]),
                    ("AT-rich 0.35/0.15", [0.35, 0.15, 0.15, 0.35])):
    u = torch.rand(n, device="cuda", generator=g)
    cdf = torch.tensor(np.cumsum(probs)[:3], device="cuda", dtype=torch.float32)
    d = letters[torch.bucketize(u, cdf)].contiguous()
    del u
    for k in ks:
        acc = c.new_accumulators(k)
        ts = []
        for it in range(5):
            lib.fkb_zero_device(ctx, k, acc.table.data_ptr(), acc.flags.data_ptr(), acc.partials.data_ptr(), st.cuda_stream)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            c.count_stream_device(d, k, acc)
            e1.record(st)
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sorted(ts[1:])[2]
        res = c.finalize_device(acc, n)
        c.set_variant(1)
        acc2 = c.new_accumulators(k)
        c.count_stream_device(d, k, acc2)
        ref = c.finalize_device(acc2, n)
        c.set_variant(0)
        ok = bool(np.array_equal(res.table, ref.table))
        print(f"{name:26s} k={k:2d} {ms:8.3f} ms {n / ms / 1e6:9.1f} Gbases/s  max count {int(res.table.max())}  equals direct kernel: {ok}", flush=True)
        del acc, acc2
    del d
    torch.cuda.empty_cache()
