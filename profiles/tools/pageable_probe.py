#!/usr/bin/env python
"""Pageable input (what the CLI's mmap'ed file is) through the host loader pipeline: host strip threads -> pinned
slots -> H2D -> count.  Run on a GPU box:   python profiles/tools/pageable_probe.py [n_bases]
Prints the strip alone and the whole C-ABI call for each host strip path (FKB_STRIP_ISA)."""
import ctypes
import os
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import numpy as np
import torch
from findkmer_b200 import synth
from findkmer_b200._lib import FkbCounts
from findkmer_b200.engine import KmerCounter

n_bases = int(sys.argv[1]) if len(sys.argv) > 1 else 3_100_000_000
k = 11
c = KmerCounter(0)
raw = c.synth_fasta_device(synth.config4(n_bases=n_bases)).cpu().numpy()  # pageable
print(f"host cores {os.cpu_count()}, raw bytes {raw.size/1e9:.3f} GB, cpu: "
      + next((l.split(':')[1].strip() for l in open('/proc/cpuinfo') if l.startswith('model name')), '?'))
flags = next((l for l in open('/proc/cpuinfo') if l.startswith('flags')), '')
print("avx512_vbmi2:", 'avx512_vbmi2' in flags)
out = np.empty(raw.size + 64, dtype=np.uint8)
out[:] = 0
h_table = torch.empty(4 ** k, dtype=torch.int32, pin_memory=True)
cnt = FkbCounts()
for isa in ("avx2", "native"):
    if isa == "native":
        os.environ.pop("FKB_STRIP_ISA", None)
    else:
        os.environ["FKB_STRIP_ISA"] = isa
    n = ctypes.c_size_t(0)
    best = 1e9
    for _ in range(2):
        t0 = time.perf_counter()
        c._lib.fkb_strip_fasta(raw.ctypes.data, raw.size, out.ctypes.data, ctypes.byref(n), 0)
        best = min(best, time.perf_counter() - t0)
    ts = []
    for i in range(4):
        t0 = time.perf_counter()
        c._check(c._lib.fkb_count_fasta_host(c._ctx, raw.ctypes.data, raw.size, k, h_table.data_ptr(), ctypes.byref(cnt)))
        ts.append((time.perf_counter() - t0) * 1e3)
    print(f"{isa:7s} strip alone (two passes + copy) {best*1e3:7.1f} ms | pageable e2e calls: "
          + " ".join(f"{t:6.1f}" for t in ts) + f" ms  best {n_bases/min(ts)/1e6:6.1f} Gbases/s  (N={cnt.n_kmers})")

os.environ.pop("FKB_STRIP_ISA", None)
for slots in (12, 24, 32, 48, 64):
    c._check(c._lib.fkb_set_option(c._ctx, b"loader_slots", slots))
    for threads in (12, 15, 16):
        os.environ["FKB_HOST_THREADS"] = str(threads)
        ts = []
        for i in range(4):
            t0 = time.perf_counter()
            c._check(c._lib.fkb_count_fasta_host(c._ctx, raw.ctypes.data, raw.size, k, h_table.data_ptr(), ctypes.byref(cnt)))
            ts.append((time.perf_counter() - t0) * 1e3)
        print(f"slots {slots:2d} threads {threads:2d}: " + " ".join(f"{t:6.1f}" for t in ts) + f" ms  (N={cnt.n_kmers})")
