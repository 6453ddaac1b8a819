#!/usr/bin/env python
"""A small end-to-end exercise of every kernel (direct, bucketed at two strides, device loader, finalize, generator) for
`compute-sanitizer --tool memcheck python profiles/tools/sanity_small.py` -- sizes kept tiny because the tool is ~50x slower."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import numpy as np
from findkmer_b200 import synth
from findkmer_b200.engine import KmerCounter
from oracle import harness

c = KmerCounter(0)
data = synth.render(synth.config5(n_bases=400_000))
fixture = (Path(__file__).resolve().parents[2] / "tests" / "golden" / "test.txt").read_bytes()
for loader in (1, 2):
    c.set_loader(loader)
    c.set_loader_chunk(150_000 if loader == 2 else 0)
    for variant in (1, 2):
        c.set_variant(variant)
        for k, d in ((6, fixture), (8, data), (11, data), (13, data)):
            got = c.count_fasta(d, k)
            want = harness.oracle_count_fasta(d, k)
            assert np.array_equal(got.table, want.table) and got.node_count == want.node_count and got.unknown_chars == want.unknown_chars, (loader, variant, k)
dev = c.synth_fasta_device(synth.config2(n_bases=100_000)).cpu().numpy()
assert np.array_equal(dev, synth.render(synth.config2(n_bases=100_000)))
c.close()
print("sanity_small ok")
