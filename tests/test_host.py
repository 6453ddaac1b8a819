"""Host-side logic and the C-ABI surface, on CPU (no compute calls that need a GPU)."""
import ctypes
import hashlib
import re
from pathlib import Path

import numpy as np
import pytest

from conftest import GOLDEN_DIR, ROOT, random_fasta


def test_library_exports_every_declared_symbol(fkb_lib):
    from findkmer_b200 import _lib
    header = (ROOT / "include" / "findkmer_b200.h").read_text()
    declared = sorted(set(re.findall(r"\b(fkb_[a-z0-9_]+)\s*\(", header)))
    assert declared, "no declarations found"
    assert sorted(_lib.EXPORTS) == declared, "findkmer_b200/_lib.py::EXPORTS is out of sync with the header"
    for name in declared:
        assert getattr(fkb_lib, name) is not None, name


def test_abi_struct_sizes_match_header(fkb_lib):
    from findkmer_b200._lib import FkbCounts, FkbPartials
    assert ctypes.sizeof(FkbCounts) == 12 * 8 and ctypes.sizeof(FkbPartials) == 12 * 8
    assert fkb_lib.fkb_table_entries(11) == 4 ** 11 and fkb_lib.fkb_table_entries(17) == 0
    assert fkb_lib.fkb_prefix_flags_bytes(11) == sum(4 ** d for d in range(1, 11))
    assert fkb_lib.fkb_max_nodes(11) == 5592405 and fkb_lib.fkb_max_nodes(6) == 5461


def test_no_gpu_means_loud_failure_not_fallback(fkb_lib):
    """without a usable sm_100 device every context creation fails; there is no CPU path behind the ABI"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from findkmer_b200 import FindKmerError, _lib
    from findkmer_b200.engine import KmerCounter
    with pytest.raises(FindKmerError) as e:
        KmerCounter(0)
    assert e.value.status == _lib.FKB_ERR_CUDA


def test_product_does_not_touch_the_oracle():
    """the oracle is test infrastructure: nothing under findkmer_b200/ may import, link or name it"""
    for path in (ROOT / "findkmer_b200").rglob("*"):
        if path.suffix in {".py", ".cu", ".cuh", ".cpp", ".h"}:
            text = path.read_text(errors="replace")
            assert "oracle" not in text.lower(), path


# ---- the host loader: stream contract ------------------------------------------------------------------
@pytest.fixture(params=["scalar", "avx2", "native"])
def strip_isa(request, monkeypatch):
    """the loader picks its widest vector path at run time (fkb_loader.cpp); FKB_STRIP_ISA lowers it per call"""
    if request.param != "native":
        monkeypatch.setenv("FKB_STRIP_ISA", request.param)
    else:
        monkeypatch.delenv("FKB_STRIP_ISA", raising=False)
    return request.param


def test_strip_dense_special_bytes(harness, strip_isa):
    """vector steps with many line breaks, '>' and 0xFF per word, and the bytes the 0x3E-bit prefilter also catches"""
    from findkmer_b200.engine import KmerCounter
    rng = np.random.default_rng(7)
    compared = 0
    alphabet = np.frombuffer(b"ACGTNacgt\n\n\n>?~\x7f\xbe\xbf\xfe\x3f", dtype=np.uint8)
    for n in (1, 63, 64, 65, 127, 128, 129, 191, 192, 193, 1000, 70_000):
        for trial in range(6):
            data = alphabet[rng.integers(0, len(alphabet), n)].copy()
            if trial in (2, 3, 5):  # sparse: long clean runs between the special bytes (the two-vector inner loop)
                data[rng.random(n) < 0.985] = ord("G")
            if trial % 2:
                data[rng.integers(0, n)] = 0xFF
            if trial >= 4:
                data[data == ord(">")] = ord("A")  # no headers: the 0xFF ends the scan wherever it is
            data[-1] = ord("\n")
            rc, want = harness.oracle_strip(data)
            if rc == harness.FKO_OK:
                assert bytes(KmerCounter.strip(data, 2)) == bytes(want), (n, trial)
                compared += 1
    assert compared >= 60


def test_strip_property_any_bytes(harness, strip_isa):
    """hypothesis: any byte string over a FASTA-flavoured alphabet strips exactly as the oracle's scan says, status included"""
    from hypothesis import given, settings, strategies as st
    from findkmer_b200 import FindKmerError, _lib
    from findkmer_b200.engine import KmerCounter
    status_of = {harness.FKO_ERR_EMPTY: _lib.FKB_ERR_EMPTY_INPUT, harness.FKO_ERR_UNTERMINATED_HDR: _lib.FKB_ERR_UNTERMINATED_HEADER}
    piece = st.one_of(st.sampled_from([b"\n", b">", b"\xff", b"ACGT" * 20, b"N" * 70, b"acgtn", b">hdr \xff >\n", b"?~\x7f\xbe"]),
                      st.binary(min_size=1, max_size=40))

    @settings(max_examples=150, deadline=None)
    @given(st.lists(piece, min_size=0, max_size=30))
    def check(pieces):
        data = b"".join(pieces)
        rc, want = harness.oracle_strip(data)
        if rc == harness.FKO_OK:
            assert bytes(KmerCounter.strip(data, 2)) == bytes(want)
        else:
            with pytest.raises(FindKmerError) as e:
                KmerCounter.strip(data, 2)
            assert e.value.status == status_of[rc]

    check()


def test_strip_matches_contract_on_golden_inputs(harness, golden, test_txt, strip_isa):
    from findkmer_b200 import FindKmerError, _lib
    from findkmer_b200.engine import KmerCounter
    inputs = [test_txt] + [r["input_latin1"].encode("latin1") for r in golden["micro"]] + [random_fasta(s, 9000) for s in range(4)]
    for data in inputs:
        rc, want = harness.oracle_strip(data)
        if rc == harness.FKO_OK:
            assert bytes(KmerCounter.strip(data)) == bytes(want)
        else:
            with pytest.raises(FindKmerError) as e:
                KmerCounter.strip(data)
            assert e.value.status == {harness.FKO_ERR_EMPTY: _lib.FKB_ERR_EMPTY_INPUT,
                                      harness.FKO_ERR_UNTERMINATED_HDR: _lib.FKB_ERR_UNTERMINATED_HEADER}[rc]


@pytest.mark.parametrize("threads", [1, 3, 8])
def test_strip_block_boundaries(harness, threads, strip_isa):
    """> 4 MiB inputs cross loader blocks: headers, newlines and 0xFF placed on and around the block edges"""
    from findkmer_b200 import synth
    from findkmer_b200.engine import KmerCounter
    data = synth.render(synth.config2(n_bases=9_500_000)).copy()
    block = 4 << 20
    for edge in (block, 2 * block):
        patch = np.frombuffer(b"ACGT>this header line straddles the loader block edge ACGT>>\nACGTNNacgt\n\n\nACGT", dtype=np.uint8)
        data[edge - 40:edge - 40 + len(patch)] = patch
    rc, want = harness.oracle_strip(data)
    assert rc == 0
    assert np.array_equal(KmerCounter.strip(data, threads), want)
    data[block + 5] = 0xFF  # inside the header that straddles the edge: skipped like any header byte
    rc, want = harness.oracle_strip(data)
    assert np.array_equal(KmerCounter.strip(data, threads), want)
    data[2 * block + 35] = 0xFF  # outside a header: ends the scan
    rc, want = harness.oracle_strip(data)
    got = KmerCounter.strip(data, threads)
    assert np.array_equal(got, want) and len(got) < 9_000_000


def test_strip_giant_lines_and_header_only(harness, strip_isa):
    from findkmer_b200.engine import KmerCounter
    one_line = b">h\n" + b"ACGT" * 3_000_000 + b"\n"
    assert np.array_equal(KmerCounter.strip(one_line, 4), harness.oracle_strip(one_line)[1])
    giant_header = b"ACGT\n>" + b"x" * 9_000_000 + b"\nACGT\n"
    assert bytes(KmerCounter.strip(giant_header, 4)) == b"ACGT>ACGT"


# ---- the writer against the reference's files ---------------------------------------------------------
def _counts_from_oracle(o, k):
    from findkmer_b200._lib import FkbCounts
    from findkmer_b200.engine import KmerCounts
    c = FkbCounts()
    c.n_kmers, c.base_total, c.node_count = o.n_kmers, o.base_total, o.node_count
    for i in range(4):
        c.base_count[i] = o.base_count[i]
    return KmerCounts.from_struct(k, o.table, c)


def test_writer_reproduces_golden_files(harness, golden, test_txt, tmp_path):
    """product writer fed the table the reference would have built => the reference's bytes"""
    from findkmer_b200.engine import write_outputs
    for k, z, name in [(6, None, "6mer_Historam_Of_test.txt.csv"), (11, None, "11mer_Historam_Of_test.txt.csv"),
                       (6, 1, "6mer_Historam_Of_test.txtzScoreFiltered_z1.csv"), (6, 2, "6mer_Historam_Of_test.txtzScoreFiltered_z2.csv")]:
        kc = _counts_from_oracle(harness.oracle_count_fasta(test_txt, k), k)
        csv, stats = tmp_path / "a.csv", tmp_path / "a.txt"
        assert write_outputs(kc, csv, stats, z_threshold=z, n_threads=3) == 0
        assert csv.read_bytes() == (GOLDEN_DIR / name).read_bytes()
        assert stats.read_bytes() == (GOLDEN_DIR / f"{k}mer_Base_Stats_Of_test.txt.txt").read_bytes()
    for k in (1, 3, 7, 8, 9, 10):
        kc = _counts_from_oracle(harness.oracle_count_fasta(test_txt, k), k)
        csv, stats = tmp_path / "b.csv", tmp_path / "b.txt"
        write_outputs(kc, csv, stats)
        assert hashlib.sha256(csv.read_bytes()).hexdigest() == golden["test_txt"][str(k)]["csv_sha256"]
        assert hashlib.sha256(stats.read_bytes()).hexdigest() == golden["test_txt"][str(k)]["stats_sha256"]


def test_writer_edge_cases_against_golden(harness, golden, tmp_path):
    """zero-probability base => header-only CSV + status; no run >= k => NaN statistics, no rows"""
    from findkmer_b200 import _lib
    from findkmer_b200.engine import write_outputs
    for rec in golden["micro"]:
        data = rec["input_latin1"].encode("latin1")
        if rec["hung"] or data == b"":
            continue
        k = rec["k"]
        kc = _counts_from_oracle(harness.oracle_count_fasta(data, k), k)
        csv, stats = tmp_path / "c.csv", tmp_path / "c.txt"
        status = write_outputs(kc, csv, stats)
        assert (status == _lib.FKB_ERR_ZERO_BASE_PROBABILITY) == rec["division_overflow"], rec
        assert hashlib.sha256(csv.read_bytes()).hexdigest() == rec["csv_sha256"], rec
        assert hashlib.sha256(stats.read_bytes()).hexdigest() == rec["stats_sha256"], rec


def test_writer_live_against_reference_binary(harness, tmp_path):
    if not harness.reference_available():
        pytest.skip("reference binary not built")
    from findkmer_b200 import synth
    from findkmer_b200.engine import write_outputs
    data = synth.render(synth.config2(n_bases=400_000))
    for k, z in ((7, None), (9, 3), (11, None)):
        r = harness.run_reference(data, k, z=z)
        kc = _counts_from_oracle(harness.oracle_count_fasta(data, k), k)
        csv, stats = tmp_path / "d.csv", tmp_path / "d.txt"
        assert write_outputs(kc, csv, stats, z_threshold=z) == 0
        assert csv.read_bytes() == r.csv and stats.read_bytes() == r.stats


def _tsv_join(path1, path2):
    """What findKmer/mergeFile4GNUPLOT.pl:13-32 is meant to do: split both files on tabs, key on column 0, and for keys in
    both emit  col0 \t col1(file1) \t col2(file1) \t col2(file2)."""
    first = {}
    for ln in open(path1, "rb").read().split(b"\n"):
        f = ln.split(b"\t")
        if len(f) >= 3:
            first[f[0]] = (f[1], f[2])
    out = {}
    for ln in open(path2, "rb").read().split(b"\n"):
        f = ln.split(b"\t")
        if len(f) >= 3 and f[0] in first:
            out[f[0]] = b"\t".join((f[0],) + first[f[0]] + (f[2],))
    return out


def test_tsv_export_feeds_the_gnuplot_join(harness, test_txt, tmp_path):
    """the histogram once more as the tab-separated table mergeFile4GNUPLOT.pl joins (kmer, h, frequency, ...).
    Two inputs (the launcher's upstream-vs-full comparison in miniature) -> two TSVs -> the join: every k-mer present in both
    comes out with both counts, and the TSV rows are the CSV rows re-ordered (same text, so the same numbers)."""
    import shutil
    import subprocess
    from findkmer_b200 import synth
    from findkmer_b200.engine import write_outputs
    k = 6
    inputs = [test_txt, synth.render(synth.config3(n_records=40)).tobytes()]
    tables = []
    for i, data in enumerate(inputs):
        o = harness.oracle_count_fasta(data, k)
        kc = _counts_from_oracle(o, k)
        csv, stats, tsv = tmp_path / f"{i}.csv", tmp_path / f"{i}.txt", tmp_path / f"{i}.tsv"
        assert write_outputs(kc, csv, stats, tsv_path=tsv, n_threads=2) == 0
        rows = csv.read_bytes().split(b"\n")[1:]
        lines = tsv.read_bytes().split(b"\n")
        assert lines[-1] == b"" and len(lines) - 1 == len(rows) == int(np.count_nonzero(o.table))
        for row, line in zip(rows, lines):
            f = row.split(b", ")
            assert line.split(b"\t") == [f[0], f[1], f[3], f[2]] + f[4:]
        tables.append(o.table)
    joined = _tsv_join(tmp_path / "0.tsv", tmp_path / "1.tsv")
    both = np.flatnonzero((tables[0] != 0) & (tables[1] != 0))
    assert len(joined) == len(both) > 100
    for code in both[:: max(1, len(both) // 50)]:
        kmer = harness.code_to_kmer(int(code), k).encode()
        f = joined[kmer].split(b"\t")
        assert (int(f[2]), int(f[3])) == (int(tables[0][code]), int(tables[1][code]))
    # where the reference tree and perl exist (this container, not the GPU box): the reference's own script with its two
    # typos repaired on the fly (`split(/\t/. $line)` -> `,`; out-file argument index 3 -> 2) gives the same table
    script = Path("/root/reference/findKmer/mergeFile4GNUPLOT.pl")
    if script.exists() and shutil.which("perl"):
        fixed = script.read_text().replace("split(/\\t/. $line)", "split(/\\t/, $line)").replace("$ARGV[3]", "$ARGV[2]")
        (tmp_path / "merge_fixed.pl").write_text(fixed)
        subprocess.run(["perl", "merge_fixed.pl", "0.tsv", "1.tsv", "merged.dat"], cwd=tmp_path, check=True, timeout=60)
        got = {ln.split(b"\t")[0]: ln for ln in (tmp_path / "merged.dat").read_bytes().split(b"\n") if ln}
        assert got == joined


def test_cli_progress_lines_of_a_non_quiet_run(harness, test_txt, tmp_path):
    """-q 0: `Read <baseCounter> bases` + the echoed header line at every record (findKmer.cpp:996-1002), byte for byte
    against the untouched reference on its own fixture.  The lines are printed before the GPU is touched, so this runs
    without one (the program then stops with its no-GPU message; on a GPU box it simply carries on)."""
    import subprocess
    exe = ROOT / "findkmer_b200" / "bin" / "findKmer"
    if not exe.exists() or not harness.reference_available():
        pytest.skip("CLI or reference binary not built")

    def segment(text):
        a = text.index("!!!Find The KMER!!!")
        b = text.find("Statistics of occurrences", a)
        return text[a:b] if b >= 0 else text[a:]

    for k in (6, 11):
        ref = harness.run_reference(test_txt, k, name="test.txt", probe=False)
        work = tmp_path / f"q0_{k}"
        work.mkdir()
        (work / "test.txt").write_bytes(test_txt)
        mine = subprocess.run([str(exe), "-q", "0", "-k", str(k), "-p", "test.txt"], cwd=work, capture_output=True, text=True,
                              stdin=subprocess.DEVNULL, timeout=300)
        quiet_ref = segment(ref.stdout)
        assert "Read " not in quiet_ref  # harness runs the reference with -q 1
        r0 = subprocess.run([str(harness.REF_BIN), "-q", "0", "-k", str(k), "-p", "test.txt"], cwd=work, capture_output=True, text=True,
                            stdin=subprocess.DEVNULL, timeout=60)
        want = segment(r0.stdout)
        assert want.count("Read ") == 6 and ">ENST00000019317" in want
        assert segment(mine.stdout) == want


# ---- synthetic inputs -----------------------------------------------------------------------------------
def test_synth_layouts_are_deterministic_and_shaped():
    from findkmer_b200 import synth
    lay = synth.config3(n_records=5)
    f = bytes(synth.render(lay))
    lines = f.split(b"\n")
    assert lines[0] == b">ENST00000000000" and len(lines[1]) == 1001 and set(lines[1]) <= set(b"ACGT")
    assert hashlib.sha256(f).hexdigest() == hashlib.sha256(bytes(synth.render(synth.config3(n_records=5)))).hexdigest()
    lay2 = synth.config2(n_bases=1000)
    f2 = bytes(synth.render(lay2))
    assert all(len(ln) <= 60 for ln in f2.split(b"\n")) and f2.count(b">") == 16 and len(f2) == lay2.total_bytes
    f5 = synth.render(synth.config5(n_bases=2_000_000))
    frac_n = np.count_nonzero(f5 == ord("N")) / 2e6
    frac_lc = np.count_nonzero((f5 >= ord("a")) & (f5 <= ord("z")) & (np.cumsum(f5 == 10) >= 0)) / 2e6
    assert 0.02 < frac_n < 0.09 and 0.05 < frac_lc < 0.2
    assert synth.config4().record_bases(0) < 2 ** 31  # the reference's seqSize is an int


def test_stripped_layout_equals_loader_output():
    from findkmer_b200 import synth
    from findkmer_b200.engine import KmerCounter
    for lay in (synth.config2(n_bases=100_001), synth.config3(n_records=40), synth.config5(n_bases=120_000)):
        assert np.array_equal(KmerCounter.strip(synth.render(lay)), synth.stripped_stream(lay))
