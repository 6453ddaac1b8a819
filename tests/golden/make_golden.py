"""Regenerate tests/golden/golden.json and the golden CSV / Base_Stats files.

Run where the reference exists (this container):  python tests/golden/make_golden.py
It executes the UNTOUCHED reference binary (oracle/_ref/findKmer_probe, built by oracle/Makefile from
/root/reference) on:
  * tests/golden/test.txt -- a byte-for-byte copy of the reference's only fixture (findKmer/test.txt,
    sha256 359e996b...f5d823), for k = 1,3,6..11,16,20
  * the known-answer micro-vectors of SURVEY.md section 4 (reset rules), k = 3 and 7
  * small seeded random inputs with junk bytes, headers, CRLF, N runs and lower case, k = 1..12
and records what it printed/wrote.  The fixtures are what the GPU box (which has no /root/reference)
checks the oracle and the CUDA path against.
"""
import hashlib
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import harness as H  # noqa: E402

OUT = Path(__file__).resolve().parent


def sha(b: bytes) -> str:
    return hashlib.sha256(b).hexdigest()


def nonzero(table):
    idx = np.flatnonzero(table)
    return [[int(i), int(table[i])] for i in idx]


def record(data: bytes, k: int, z=None, name="in.fa", timeout=20.0):
    r = H.run_reference(data, k, z=z, name=name, timeout=timeout)
    rec = {"k": k, "z": z, "ok": r.ok, "hung": r.hung, "exit_code": r.exit_code, "csv_sha256": sha(r.csv), "stats_sha256": sha(r.stats),
           "node_count": r.node_count, "base_count": list(r.base_count) if r.base_count else None, "base_total": r.base_total,
           "division_overflow": "Division overflow detected" in r.stdout, "empty_file": r.exit_code == 1 and not r.stdout.count("Statistics")}
    if r.ok and z is None and k <= 16:
        t = H.csv_to_table(r.csv, k)
        rec["n_kmers"] = int(t.sum(dtype=np.uint64))
        rec["distinct"] = int(np.count_nonzero(t))
    return rec, r


def random_input(seed: int, n: int) -> bytes:
    rng = np.random.default_rng(seed)
    alphabet = np.frombuffer(b"ACGT" * 12 + b"NNacgt\r \tRY-*.0" + b"\n\n\n", dtype=np.uint8)
    body = alphabet[rng.integers(0, alphabet.size, size=n)].copy()
    # sprinkle header lines (always newline-terminated) and make sure all four bases occur in long runs
    out = bytearray(b">seed%d first header\n" % seed)
    pos = 0
    while pos < n:
        step = int(rng.integers(20, 400))
        out += bytes(body[pos:pos + step])
        pos += step
        if rng.random() < 0.3:
            out += b">hdr with ACGT and > inside %d\n" % pos
    out += b"\nACGTACGTTGCATGCAAACCGGTTACGTACGTTGCATGCAAACCGGTT\n"
    return bytes(out)


def main():
    golden = {"source": "oracle/_ref/findKmer_probe (untouched reference, g++ -O3 -w)", "test_txt": {}, "micro": [], "random": []}
    test_txt = (OUT / "test.txt").read_bytes()
    golden["test_txt_sha256"] = sha(test_txt)
    for k in (1, 3, 6, 7, 8, 9, 10, 11, 16, 20):
        rec, r = record(test_txt, k, name="test.txt")
        golden["test_txt"][str(k)] = rec
        if k in (6, 11):
            (OUT / f"{k}mer_Historam_Of_test.txt.csv").write_bytes(r.csv)
            (OUT / f"{k}mer_Base_Stats_Of_test.txt.txt").write_bytes(r.stats)
    for z in (1, 2):
        rec, r = record(test_txt, 6, z=z, name="test.txt")
        golden["test_txt"][f"6_z{z}"] = rec
        (OUT / f"6mer_Historam_Of_test.txtzScoreFiltered_z{z}.csv").write_bytes(r.csv)

    micro_inputs = [
        b"ACGTACGTAC\n", b"ACGT\nACGT\nACG\n", b"ACGTacgtACGTAC\n", b"ACGT\r\nACGT\r\nACG\n", b"ACGTA>hdr ACGT\nCGTAC\n",
        b"ACGTAC\xffGTACGT\n", b"ACGT ACGT\tACGT\n", b"ACGTNACGTRACGT-ACGT*ACGT\n", b"ACGTACGTAC", b"ACGTACGT\n>hdr_no_newline",
        b"ACACACACAC\n", b"", b"AC\nGT\nAC\n", b">only a header\n", b"\n\n\n", b"N", b"A", b"ACGT", b">h\xff\nACGTACGT\xffACGT\n",
        b"ANCNGNTNACGNACGTN\nACGT", b">a\n>b\nAC>c\nGT\nACGTT\n",
    ]
    for data in micro_inputs:
        for k in (3, 7):
            rec, r = record(data, k, timeout=3.0)
            rec["input_latin1"] = data.decode("latin1")
            if r.ok and k <= 16:
                rec["table_nonzero"] = nonzero(H.csv_to_table(r.csv, k))
            golden["micro"].append(rec)
    for seed in range(6):
        data = random_input(seed, 6000 + 1500 * seed)
        for k in (1, 2, 5, 8, 12):
            rec, r = record(data, k)
            rec["seed"] = seed
            rec["n"] = 6000 + 1500 * seed
            rec["input_sha256"] = sha(data)
            rec["table_sha256"] = sha(H.csv_to_table(r.csv, k).tobytes()) if r.ok else None
            golden["random"].append(rec)
    (OUT / "golden.json").write_text(json.dumps(golden, indent=1, sort_keys=True) + "\n")
    print("wrote", OUT / "golden.json")


if __name__ == "__main__":
    main()
