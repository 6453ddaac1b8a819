"""Full-size golden records for BASELINE.json configs 4 and 5 (3.1e9 bases): tests/golden/golden_fullsize.json.

    python tests/golden/make_golden_fullsize.py            # ~5 minutes on 8 cores, ~6 GB of RAM

The untouched reference binary needs ~30 minutes per run at this size and cannot be sharded, so the generator is the
C restatement (oracle/kmer_oracle.c -- pinned against the binary by tests/test_oracle.py) run RECORD-PARALLEL: every
record starts with a header line, i.e. a window reset (findKmer.cpp:994), so no window, run or short-run prefix
spans two records and

    table, n_kmers, base_total, base_count[4], unknown_chars   are sums over records,
    node_count                                                 is the size of a UNION of per-record prefix sets.

The union is not derivable from per-record node counts in general.  It is here: the script checks that every one of
the 4^k k-mers occurs in the summed table, so every trie node exists and node_count = 1 + sum_{d<=k} 4^d
(estimate_RAM_usage, findKmer.cpp:1256-1260); short-run prefixes (:1059-1062) can add nothing to a full trie.

What is recorded per (config, k): sha256 of the little-endian uint32 table, the scalars, and a few spot values.
The GPU tests (tests/test_gpu_parity.py::test_full_size_against_golden) count the same synthetic input -- generated
on the device by the bit-identical CUDA twin of findkmer_b200/synth.py -- and compare.
"""
from __future__ import annotations

import hashlib
import json
import sys
import time
from multiprocessing import Pool
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

OUT = Path(__file__).resolve().parent / "golden_fullsize.json"
CHUNK = 8_000_000  # bases rendered per numpy call (bounds the temporaries)


def _layout(name: str):
    from findkmer_b200 import synth
    return {"config4": synth.config4, "config5": synth.config5}[name]()


def _record_stream(lay, r: int) -> np.ndarray:
    """Record r in the stripped layout: one '>' byte, then its bases (what the loader hands to the scan)."""
    from findkmer_b200 import synth
    _, base0 = lay.record_offsets()
    nb = lay.record_bases(r)
    out = np.empty(nb + 1, dtype=np.uint8)
    out[0] = ord(">")
    g0 = int(base0[r])
    for a in range(0, nb, CHUNK):
        b = min(nb, a + CHUNK)
        out[1 + a:1 + b] = synth.base_letters(lay, np.arange(g0 + a, g0 + b, dtype=np.uint64))
    return out


def _work(job):
    name, ks, r = job
    from oracle import harness as H
    lay = _layout(name)
    stream = _record_stream(lay, r)
    res = {}
    for k in ks:
        o = H.oracle_count_stream(stream, k)
        assert o.rc == 0, (name, k, r, o.rc)
        res[k] = (o.table, o.n_kmers, o.base_total, tuple(int(x) for x in o.base_count), o.unknown_chars)
    return r, res


def main():
    jobs_spec = [("config4", [11, 8]), ("config5", [8, 11])]
    doc = {"generator": "tests/golden/make_golden_fullsize.py (oracle/kmer_oracle.c, record-parallel)", "records": {}}
    t0 = time.time()
    with Pool(8) as pool:
        for name, ks in jobs_spec:
            lay = _layout(name)
            tables = {k: np.zeros(4 ** k, dtype=np.uint64) for k in ks}
            scal = {k: {"n_kmers": 0, "base_total": 0, "base_count": [0, 0, 0, 0], "unknown_chars": 0} for k in ks}
            for r, res in pool.imap_unordered(_work, [(name, ks, r) for r in range(lay.n_records)]):
                for k, (t, n, bt, bc, unk) in res.items():
                    tables[k] += t
                    s = scal[k]
                    s["n_kmers"] += int(n)
                    s["base_total"] += int(bt)
                    s["unknown_chars"] += int(unk)
                    for i in range(4):
                        s["base_count"][i] += bc[i]
                print(f"{name} record {r} done ({time.time() - t0:.0f} s)", flush=True)
            for k in ks:
                t = tables[k]
                assert int(t.max()) < 2 ** 32
                assert int(t.min()) > 0, "a k-mer is missing: node_count is not the full trie"
                t32 = t.astype("<u4")
                rec = dict(scal[k])
                rec.update({"config": name, "k": k, "n_bases": lay.n_bases, "n_records": lay.n_records, "seed": lay.seed,
                            "table_sha256": hashlib.sha256(t32.tobytes()).hexdigest(),
                            "node_count": 1 + sum(4 ** d for d in range(1, k + 1)),
                            "table_sum": int(t.sum()), "table_max": int(t.max()), "table_min": int(t.min()),
                            "spot": {str(i): int(t[i]) for i in (0, 1, 4 ** k // 3, 4 ** k - 1)}})
                assert rec["table_sum"] == rec["n_kmers"]
                doc["records"][f"{name}_k{k}"] = rec
                print(json.dumps(rec), flush=True)
    OUT.write_text(json.dumps(doc, indent=1) + "\n")
    print(f"wrote {OUT} in {time.time() - t0:.0f} s")


if __name__ == "__main__":
    main()
