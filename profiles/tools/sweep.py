#!/usr/bin/env python
"""Device-resident count throughput for the BASELINE.json shapes and k = 6..11 (not the bench line: evidence for DESIGN.md).
    python profiles/tools/sweep.py [shape-substring [k,k,... [n_bases [variant,variant,...]]]] > profiles/rNN_sweep.txt
variants: 0 auto, 1 direct, 2 bucketed (13-mers), 3 single-pass shared memory (k <= 8), 4 bucketed 16-mer items (k = 11)"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import torch
from findkmer_b200 import synth
from findkmer_b200.engine import KmerCounter

c = KmerCounter(0)
lib, ctx = c._lib, c._ctx
shapes = [("config2 12 Mbp, 16 records", synth.config2()), ("config3 6000 x 1001 bp", synth.config3()),
          ("config4 3.1 Gbp", synth.config4()), ("config5 3.1 Gbp, N runs + soft mask", synth.config5())]
only_shape = sys.argv[1] if len(sys.argv) > 1 else ""
only_k = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [6, 7, 8, 9, 10, 11]
n_override = int(sys.argv[3]) if len(sys.argv) > 3 else 0
print(f"{'shape':40s} {'k':>3s} {'variant':>8s} {'ms':>9s} {'Gbases/s':>10s}")
for name, lay in shapes:
    if only_shape not in name:
        continue
    if n_override:
        from dataclasses import replace
        lay = replace(lay, n_bases=n_override)
    d = c.synth_fasta_device(lay.stripped())
    n = d.numel()
    st = torch.cuda.current_stream()
    for k in only_k:
        acc = c.new_accumulators(k)
        variants = [int(x) for x in sys.argv[4].split(",")] if len(sys.argv) > 4 else ((0,) if n > (1 << 28) else (1, 2))
        for variant in variants:
            if (variant == 3 and k > 8) or (variant == 4 and k != 11):
                continue
            c.set_variant(variant)
            times = []
            for it in range(6):
                lib.fkb_zero_device(ctx, k, acc.table.data_ptr(), acc.flags.data_ptr(), acc.partials.data_ptr(), st.cuda_stream)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(st)
                c.count_stream_device(d, k, acc)
                e1.record(st)
                torch.cuda.synchronize()
                times.append(e0.elapsed_time(e1))
            ms = sorted(times[2:])[len(times[2:]) // 2]
            print(f"{name:40s} {k:3d} {['auto', 'direct', 'bucketed', 'smem', 'bucket16'][variant]:>8s} {ms:9.3f} {lay.n_bases / ms / 1e6:10.1f}")
        c.set_variant(0)
        del acc
    del d
    torch.cuda.empty_cache()
