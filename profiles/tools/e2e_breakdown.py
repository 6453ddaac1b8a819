#!/usr/bin/env python
"""Where does the end-to-end (host FASTA bytes -> counts on host) time go?  Run on a GPU box:
    python profiles/tools/e2e_breakdown.py [n_bases]
Prints: host strip alone (n threads), raw H2D copy alone, the pipelined C-ABI call, and the device-only count."""
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
lay = synth.config4(n_bases=n_bases)
d_raw = c.synth_fasta_device(lay)
h_raw = torch.empty(d_raw.numel(), dtype=torch.uint8, pin_memory=True)
h_raw.copy_(d_raw)
torch.cuda.synchronize()
raw = h_raw.numpy()
print(f"host cores {os.cpu_count()}, raw bytes {raw.size/1e9:.3f} GB")

out = np.empty(raw.size, dtype=np.uint8)
for threads in (1, 4, 8, 16, 32):
    n = ctypes.c_size_t(0)
    t0 = time.perf_counter()
    c._lib.fkb_strip_fasta(raw.ctypes.data, raw.size, out.ctypes.data, ctypes.byref(n), threads)
    dt = time.perf_counter() - t0
    print(f"strip alone, {threads:2d} threads: {dt*1e3:8.1f} ms  {raw.size/dt/1e9:6.1f} GB/s")

d = torch.empty(raw.size, dtype=torch.uint8, device="cuda")
for _ in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); d.copy_(h_raw, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"H2D alone (pinned, one copy): {dt*1e3:8.1f} ms  {raw.size/dt/1e9:6.1f} GB/s")
h_table = torch.empty(4 ** k, dtype=torch.int32, pin_memory=True)
cnt = FkbCounts()
for i in range(3):
    t0 = time.perf_counter()
    c._check(c._lib.fkb_count_fasta_host(c._ctx, h_raw.data_ptr(), h_raw.numel(), k, h_table.data_ptr(), ctypes.byref(cnt)))
    dt = time.perf_counter() - t0
    print(f"pipelined e2e call #{i}: {dt*1e3:8.1f} ms  {n_bases/dt/1e9:6.1f} Gbases/s  (N={cnt.n_kmers})")
for threads in (4, 8, 12):
    os.environ["FKB_HOST_THREADS"] = str(threads)
    t0 = time.perf_counter()
    c._check(c._lib.fkb_count_fasta_host(c._ctx, h_raw.data_ptr(), h_raw.numel(), k, h_table.data_ptr(), ctypes.byref(cnt)))
    dt = time.perf_counter() - t0
    print(f"pipelined e2e, FKB_HOST_THREADS={threads}: {dt*1e3:8.1f} ms")
