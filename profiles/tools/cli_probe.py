#!/usr/bin/env python
"""Whole-program timing of the drop-in `findKmer` binary on a genome-sized FASTA file (GPU box):
    python profiles/tools/cli_probe.py [n_bases]
Writes the synthetic file to /dev/shm (page cache, like a warm file), then runs the program as the reference's launcher does
(-q 1 -k K -z 1000 -p file) for k = 11 and 6, and once with -K 6-11; prints wall seconds of each process."""
import os
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from findkmer_b200 import synth
from findkmer_b200.engine import KmerCounter

n_bases = int(sys.argv[1]) if len(sys.argv) > 1 else 3_100_000_000
work = Path("/dev/shm/fkb_cli_probe")
work.mkdir(exist_ok=True)
fa = work / "synthetic_genome.fa"
c = KmerCounter(0)
c.synth_fasta_device(synth.config4(n_bases=n_bases)).cpu().numpy().tofile(fa)
del c
print(f"file {fa.stat().st_size/1e9:.3f} GB, {n_bases/1e9:.2f} Gbases, host cores {os.cpu_count()}")
exe = ROOT / "findkmer_b200" / "bin" / "findKmer"
runs = [(["-k", "11"], {}), (["-k", "11"], {}), (["-k", "11"], {"FKB_LOADER_SLOTS": "12"}), (["-k", "6"], {}), (["-K", "6-11"], {})]
for args, extra_env in runs:
    t0 = time.perf_counter()
    r = subprocess.run([str(exe), "-q", "1", *args, "-z", "1000", "-p", fa.name], cwd=work, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                       env=dict(os.environ, FKB_TIMING="1", **extra_env), text=True)
    dt = time.perf_counter() - t0
    print("".join(l + "\n" for l in r.stdout.splitlines() if l.startswith("[timing]")), end="")
    outs = sorted(p.name + f" ({p.stat().st_size/1e6:.1f} MB)" for p in work.iterdir() if p.name != fa.name)
    print(f"{' '.join(k + '=' + v for k, v in extra_env.items())} findKmer -q 1 {' '.join(args)} -z 1000 -p {fa.name}: rc={r.returncode} {dt:6.2f} s wall; files: {len(outs)}")
    for p in work.iterdir():
        if p.name != fa.name:
            p.unlink()
fa.unlink()
