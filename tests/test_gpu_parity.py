"""Parity of the CUDA path (through the C ABI) against the oracle and the committed golden vectors.
Every test here needs a real B200: run with `pytest -m gpu`."""
import hashlib
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR, assert_counts_equal, random_fasta

pytestmark = pytest.mark.gpu


def _device_count(counter, stream_np, k, begin=0, end=None, pieces=1):
    """device-resident path: upload a stripped stream, count [begin,end) in `pieces` launches, finalize"""
    import torch
    n = len(stream_np)
    d = torch.zeros(n + 64, dtype=torch.uint8, device="cuda")
    d[:n] = torch.from_numpy(np.ascontiguousarray(stream_np))
    acc = counter.new_accumulators(k)
    end = n if end is None else end
    cuts = np.linspace(begin, end, pieces + 1).astype(np.int64)
    for a, b in zip(cuts[:-1], cuts[1:]):
        counter.count_stream_device(d[:n], k, acc, int(a), int(b))
    return counter.finalize_device(acc, n), acc


# ---- config 1: the reference's own fixture ---------------------------------------------------------
@pytest.mark.parametrize("k", [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14])
def test_test_txt_all_k(counter, harness, test_txt, k):
    want = harness.oracle_count_fasta(test_txt, k)
    got = counter.count_fasta(test_txt, k)
    assert_counts_equal(got, want)


def test_k16_device_path_against_golden(counter, golden, test_txt):
    """k = 16 (the full 32-bit rolling register; 16 GiB table stays in HBM): scalars and sparsity vs the reference run"""
    import torch
    g = golden["test_txt"]["16"]
    stream = counter.strip(test_txt)
    d = torch.zeros(len(stream) + 64, dtype=torch.uint8, device="cuda")
    d[:len(stream)] = torch.from_numpy(stream.copy())
    acc = counter.new_accumulators(16)
    counter.count_stream_device(d[:len(stream)], 16, acc)
    got = counter.finalize_device(acc, len(stream), fetch_table=False)
    assert (got.n_kmers, list(got.base_count), got.base_total, got.node_count) == (g["n_kmers"], g["base_count"], g["base_total"], g["node_count"])
    assert int(torch.count_nonzero(acc.table)) == g["distinct"] and int(acc.table.sum(dtype=torch.int64)) == g["n_kmers"]
    del acc, d
    torch.cuda.empty_cache()


def test_test_txt_golden_numbers(counter, golden, test_txt):
    for key, g in golden["test_txt"].items():
        if "_z" in key or int(key) > 14:
            continue
        got = counter.count_fasta(test_txt, int(key))
        assert (got.n_kmers, int(np.count_nonzero(got.table)), list(got.base_count), got.base_total, got.node_count) == \
               (g["n_kmers"], g["distinct"], g["base_count"], g["base_total"], g["node_count"])


def test_csv_and_stats_bytes_from_gpu_counts(counter, golden, test_txt, tmp_path):
    """file -> GPU counts -> host writer must reproduce the reference's files byte for byte"""
    from findkmer_b200.engine import write_outputs
    for k, z, name in [(6, None, "6mer_Historam_Of_test.txt.csv"), (11, None, "11mer_Historam_Of_test.txt.csv"),
                       (6, 1, "6mer_Historam_Of_test.txtzScoreFiltered_z1.csv"), (6, 2, "6mer_Historam_Of_test.txtzScoreFiltered_z2.csv")]:
        got = counter.count_fasta(test_txt, k)
        csv, stats = tmp_path / "out.csv", tmp_path / "out.txt"
        assert write_outputs(got, csv, stats, z_threshold=z) == 0
        assert csv.read_bytes() == (GOLDEN_DIR / name).read_bytes()
        assert stats.read_bytes() == (GOLDEN_DIR / f"{k}mer_Base_Stats_Of_test.txt.txt").read_bytes()
    for k in (1, 3, 7, 8, 9, 10):
        got = counter.count_fasta(test_txt, k)
        csv, stats = tmp_path / "o.csv", tmp_path / "o.txt"
        write_outputs(got, csv, stats)
        assert hashlib.sha256(csv.read_bytes()).hexdigest() == golden["test_txt"][str(k)]["csv_sha256"]
        assert hashlib.sha256(stats.read_bytes()).hexdigest() == golden["test_txt"][str(k)]["stats_sha256"]


# ---- reset rules ---------------------------------------------------------------------------------
def test_micro_vectors(counter, harness, golden):
    from findkmer_b200 import FindKmerError, _lib
    for rec in golden["micro"]:
        data = rec["input_latin1"].encode("latin1")
        k = rec["k"]
        if rec["hung"]:
            with pytest.raises(FindKmerError) as e:
                counter.count_fasta(data, k)
            assert e.value.status == _lib.FKB_ERR_UNTERMINATED_HEADER
            continue
        if data == b"":
            with pytest.raises(FindKmerError) as e:
                counter.count_fasta(data, k)
            assert e.value.status == _lib.FKB_ERR_EMPTY_INPUT
            continue
        got = counter.count_fasta(data, k)
        assert got.node_count == rec["node_count"], rec
        if rec["ok"]:
            assert [[int(i), int(got.table[i])] for i in np.flatnonzero(got.table)] == rec["table_nonzero"], rec
            assert got.base_total == rec["base_total"] and list(got.base_count) == rec["base_count"], rec
        assert_counts_equal(got, harness.oracle_count_fasta(data, k))


@pytest.mark.parametrize("seed", range(6))
def test_random_junk_inputs(counter, harness, golden, seed):
    recs = [r for r in golden["random"] if r["seed"] == seed]
    data = random_fasta(seed, recs[0]["n"])
    for rec in recs:
        got = counter.count_fasta(data, rec["k"])
        assert hashlib.sha256(got.table.tobytes()).hexdigest() == rec["table_sha256"]
        assert (got.n_kmers, got.node_count, list(got.base_count), got.base_total) == \
               (rec["n_kmers"], rec["node_count"], rec["base_count"], rec["base_total"])
    for k in (4, 11, 13):
        assert_counts_equal(counter.count_fasta(data, k), harness.oracle_count_fasta(data, k))


def test_ragged_sizes_and_alignments(counter, harness):
    """every stream length 0..200 and every begin offset: chunk edges, halos, tails"""
    rng = np.random.default_rng(5)
    base = np.frombuffer(b"ACGTACGTACGTNACGT>acgt", dtype=np.uint8)[rng.integers(0, 22, size=400)]
    for n in list(range(0, 70)) + [127, 128, 129, 191, 192, 193, 255, 256, 257, 400]:
        s = base[:n]
        for k in (1, 5, 12):
            want = harness.oracle_count_stream(s, k)
            got = counter.count_stream(s, k)
            assert_counts_equal(got, want)


def test_device_ranges_compose(counter, harness):
    """sharding rule: a window belongs to the range that owns its LAST byte; ranges add up exactly"""
    data = random_fasta(11, 50000)
    stream = counter.strip(data)
    for k in (3, 11, 13):
        want = harness.oracle_count_stream(stream, k)
        for pieces in (1, 2, 7, 64):
            got, _ = _device_count(counter, stream, k, pieces=pieces)
            assert_counts_equal(got, want)
    # two "ranks" on one GPU: separate accumulators, then the reduction the NCCL path performs
    import torch
    k = 9
    n = len(stream)
    mid = n // 2 + 3
    _, acc0 = _device_count(counter, stream, k, 0, mid)
    _, acc1 = _device_count(counter, stream, k, mid, n)
    acc0.buf += acc1.buf            # what the fused NCCL sum-reduce does: table words and flag bytes together
    acc0.partials += acc1.partials
    assert_counts_equal(counter.finalize_device(acc0, n), harness.oracle_count_stream(stream, k))


# ---- configs 2, 3 and a slice of 5: k = 6..11 ---------------------------------------------------------
@pytest.mark.parametrize("k", [6, 7, 8, 9, 10, 11])
def test_config2_yeast_scale(counter, harness, k):
    from findkmer_b200 import synth
    data = synth.render(synth.config2())
    want = harness.oracle_count_fasta(data, k)
    assert_counts_equal(counter.count_fasta(data, k), want)


@pytest.mark.parametrize("k", [6, 7, 8, 9, 10, 11])
def test_config3_upstream_records(counter, harness, k):
    from findkmer_b200 import synth
    data = synth.render(synth.config3())
    assert_counts_equal(counter.count_fasta(data, k), harness.oracle_count_fasta(data, k))


@pytest.mark.parametrize("k", [8, 11])
def test_config5_shape_n_runs_and_soft_mask(counter, harness, k):
    from findkmer_b200 import synth
    data = synth.render(synth.config5(n_bases=6_000_000))
    assert_counts_equal(counter.count_fasta(data, k), harness.oracle_count_fasta(data, k))


def test_reference_binary_live(counter, harness):
    """the untouched reference binary travels to the GPU box prebuilt: compare against it directly"""
    if not harness.reference_available():
        pytest.skip("oracle/_ref/findKmer not present")
    from findkmer_b200 import synth
    data = synth.render(synth.config2(n_bases=600_000))
    for k in (6, 11):
        r = harness.run_reference(data, k)
        got = counter.count_fasta(data, k)
        assert r.ok
        assert np.array_equal(harness.csv_to_table(r.csv, k), got.table)
        assert (r.node_count, r.base_count, r.base_total) == (got.node_count, tuple(got.base_count), got.base_total)


# ---- the bucketed kernels (W-mers at stride S through shared memory, then folded) -------------------------
@pytest.fixture()
def bucketed(counter):
    counter.set_variant(2)
    yield counter
    counter.set_variant(0)


@pytest.mark.parametrize("k", [6, 7, 8, 9, 10, 11, 12, 13])
def test_bucketed_variant_matches_oracle(bucketed, harness, k):
    from findkmer_b200 import synth
    stream = bucketed.strip(synth.render(synth.config5(n_bases=1_500_000)))   # N runs + soft-masked runs: many resets
    want = harness.oracle_count_stream(stream, k)
    launches0 = bucketed.launches
    got, _ = _device_count(bucketed, stream, k)
    assert bucketed.launches - launches0 >= 6  # edges + bucketize + count_buckets + fold level(s) + finalize: the bucketed path really ran
    assert_counts_equal(got, want)
    got, _ = _device_count(bucketed, stream, k, 12345, len(stream) - 777, pieces=3)  # unaligned sub-ranges, three launches
    sub = harness.oracle_count_stream(stream[:len(stream) - 777], k)
    head = harness.oracle_count_stream(stream[:12345], k)
    assert np.array_equal(got.table, sub.table - head.table)


@pytest.mark.parametrize("k", [6, 11, 13])
def test_bucketed_variant_short_records(bucketed, harness, k):
    from findkmer_b200 import synth
    stream = bucketed.strip(synth.render(synth.config3(n_records=1500)))
    got, _ = _device_count(bucketed, stream, k)
    assert_counts_equal(got, harness.oracle_count_stream(stream, k))


@pytest.mark.parametrize("k", [8, 9, 11, 13])
def test_bucketed_variant_skewed_input_exercises_every_escape(bucketed, harness, k):
    """poly-A with sparse substitutions: one bucket takes almost everything -> staging rows overflow, the bucket's
    region in HBM overflows, and 16-bit counters drain at 0x8000; every escape must stay exact"""
    rng = np.random.default_rng(3)
    s = np.full(2_000_000, ord("A"), dtype=np.uint8)
    idx = rng.integers(0, s.size, size=4000)
    s[idx] = np.frombuffer(b"CGTN", dtype=np.uint8)[rng.integers(0, 4, size=idx.size)]
    got, _ = _device_count(bucketed, s, k)
    want = harness.oracle_count_stream(s, k)
    assert int(want.table.max()) > 100_000
    assert_counts_equal(got, want)


@pytest.mark.parametrize("k", [8, 11])
def test_bucketed_variant_skewed_composition_uses_the_back_region(bucketed, harness, k):
    """A/T-rich i.i.d. sequence (0.35/0.15): the A/T-rich buckets are ~5x as popular as the average, their staging rows fill
    up every tile and the surplus is appended straight to the back part of the bucket's region (and pass 2 takes the
    buckets largest first); everything must stay exact"""
    rng = np.random.default_rng(11)
    s = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.choice(4, size=6_000_000, p=[0.35, 0.15, 0.15, 0.35])]
    s = s.copy()
    s[rng.integers(0, s.size, size=300)] = ord("N")
    got, _ = _device_count(bucketed, s, k)
    assert_counts_equal(got, harness.oracle_count_stream(s, k))


# ---- k = 11: 16-mer items counted as two 13-mers in 8-bit counters (fkb_bucket2.cu) ----------------------------
@pytest.fixture(params=[0, 1], ids=["sync_flush", "ring_flush"])
def bucket16(counter, request):
    """both forms of pass 1: the default (synchronous flush) and option "p1_ring" (ring rows flushed asynchronously by their owner warps)"""
    counter.set_variant(4)
    counter._check(counter._lib.fkb_set_option(counter._ctx, b"p1_ring", request.param))
    yield counter
    counter._check(counter._lib.fkb_set_option(counter._ctx, b"p1_ring", 0))
    counter.set_variant(0)


def test_bucket16_variant_matches_oracle(bucket16, harness):
    """config-5 shape (many resets: the general path, left-over 11-mers, per-run events), whole and as three unaligned pieces;
    clean sequence (the predicate-free path); 1001-base records; seeded junk"""
    from findkmer_b200 import synth
    k = 11
    stream = bucket16.strip(synth.render(synth.config5(n_bases=1_500_000)))
    want = harness.oracle_count_stream(stream, k)
    launches0 = bucket16.launches
    got, _ = _device_count(bucket16, stream, k)
    assert bucket16.launches - launches0 >= 5  # edge slivers + bucketize16 + count_buckets16 + finalize levels
    assert_counts_equal(got, want)
    got, _ = _device_count(bucket16, stream, k, 12345, len(stream) - 777, pieces=3)
    sub = harness.oracle_count_stream(stream[:len(stream) - 777], k)
    head = harness.oracle_count_stream(stream[:12345], k)
    assert np.array_equal(got.table, sub.table - head.table)
    for lay in (synth.config2(n_bases=2_400_000), synth.config3(n_records=1500)):
        stream = bucket16.strip(synth.render(lay))
        got, _ = _device_count(bucket16, stream, k)
        assert_counts_equal(got, harness.oracle_count_stream(stream, k))
    for seed in range(3):
        stream = bucket16.strip(random_fasta(seed, 400_000))
        got, _ = _device_count(bucket16, stream, k)
        assert_counts_equal(got, harness.oracle_count_stream(stream, k))


def test_bucket16_variant_skewed_inputs_stay_exact(bucket16, harness):
    """(1) poly-A with sparse substitutions: one bucket takes almost everything -- staging rows overflow, the bucket's region
    in HBM overflows, and the 8-bit counters of pass 2 wrap, so the bucket is recounted exactly; (2) A/T-rich i.i.d. sequence:
    hot buckets use the back region and hot counters drain at 0x80 into the per-bucket list; (3) a tandem repeat of period 7."""
    k = 11
    rng = np.random.default_rng(3)
    s = np.full(3_000_000, ord("A"), dtype=np.uint8)
    idx = rng.integers(0, s.size, size=6000)
    s[idx] = np.frombuffer(b"CGTN", dtype=np.uint8)[rng.integers(0, 4, size=idx.size)]
    got, _ = _device_count(bucket16, s, k)
    want = harness.oracle_count_stream(s, k)
    assert int(want.table.max()) > 100_000
    assert_counts_equal(got, want)
    rng = np.random.default_rng(11)
    s = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.choice(4, size=8_000_000, p=[0.4, 0.1, 0.1, 0.4])].copy()
    s[rng.integers(0, s.size, size=300)] = ord("N")
    got, _ = _device_count(bucket16, s, k)
    want = harness.oracle_count_stream(s, k)
    assert int(want.table.max()) > 300          # counters beyond 0x80: the drain list is exercised
    assert_counts_equal(got, want)
    s = np.tile(np.frombuffer(b"ACGGTCA", dtype=np.uint8), 400_000)
    s[rng.integers(0, s.size, size=500)] = ord("T")
    got, _ = _device_count(bucket16, s, k)
    assert_counts_equal(got, harness.oracle_count_stream(s, k))


# ---- k <= 8: the single-pass shared-memory path (fkb_smallk.cu) ------------------------------------------------
@pytest.fixture()
def smem_path(counter):
    counter.set_variant(3)
    yield counter
    counter.set_variant(0)


@pytest.mark.parametrize("k", [1, 2, 3, 4, 5, 6, 7, 8])
def test_smem_variant_matches_oracle(smem_path, harness, k):
    """config-5 shape (N runs + soft-masked runs: the general path with its per-run events), whole and as three unaligned
    pieces; then clean sequence (the predicate-free path) and 1001-base records (a reset every kilobyte)"""
    from findkmer_b200 import synth
    stream = smem_path.strip(synth.render(synth.config5(n_bases=1_500_000)))
    want = harness.oracle_count_stream(stream, k)
    got, _ = _device_count(smem_path, stream, k)
    assert_counts_equal(got, want)
    got, _ = _device_count(smem_path, stream, k, 12345, len(stream) - 777, pieces=3)
    sub = harness.oracle_count_stream(stream[:len(stream) - 777], k)
    head = harness.oracle_count_stream(stream[:12345], k)
    assert np.array_equal(got.table, sub.table - head.table)
    for lay in (synth.config2(n_bases=1_200_000), synth.config3(n_records=1500)):
        stream = smem_path.strip(synth.render(lay))
        got, _ = _device_count(smem_path, stream, k)
        assert_counts_equal(got, harness.oracle_count_stream(stream, k))


@pytest.mark.parametrize("k", [6, 8])
def test_smem_variant_junk_and_hot_counters(smem_path, harness, k):
    """junk bytes of every kind (unknown-character count included), and poly-A with sparse substitutions: one counter takes
    almost everything, so the 16-bit counters of k = 8 drain at 0x8000 many times; both must stay exact"""
    for seed in range(3):
        data = random_fasta(seed, 300_000)
        stream = smem_path.strip(data)
        got, _ = _device_count(smem_path, stream, k)
        assert_counts_equal(got, harness.oracle_count_stream(stream, k))
    rng = np.random.default_rng(3)
    s = np.full(6_000_000, ord("A"), dtype=np.uint8)
    idx = rng.integers(0, s.size, size=4000)
    s[idx] = np.frombuffer(b"CGTN", dtype=np.uint8)[rng.integers(0, 4, size=idx.size)]
    got, _ = _device_count(smem_path, s, k)
    want = harness.oracle_count_stream(s, k)
    assert int(want.table.max()) > 1_000_000
    assert_counts_equal(got, want)


def test_auto_takes_the_smem_path_for_small_k(counter, harness):
    """variant 0 (what users get): k <= 8 on a few megabases runs count_smem_kernel, not the bucketed kernels
    (3 launches: the kernel, the edge slivers, ... plus finalize)"""
    from findkmer_b200 import synth
    stream = counter.strip(synth.render(synth.config5(n_bases=3_000_000)))
    for k in (6, 8):
        n0 = counter.launches
        got, _ = _device_count(counter, stream, k)
        assert 3 <= counter.launches - n0 <= 6
        assert_counts_equal(got, harness.oracle_count_stream(stream, k))


def test_phase_times_of_the_bucketed_path(bucketed):
    """option phase_events: the library records CUDA events around pass 1 / pass 2 / fold; results unchanged"""
    import ctypes
    from findkmer_b200 import synth
    stream = bucketed.strip(synth.render(synth.config2(n_bases=3_000_000)))
    lib, ctx = bucketed._lib, bucketed._ctx
    ms = (ctypes.c_double * 3)()
    assert lib.fkb_phase_times(ctx, ctypes.byref(ms)) != 0  # off by default
    want, _ = _device_count(bucketed, stream, 8)
    assert lib.fkb_set_option(ctx, b"phase_events", 1) == 0
    try:
        got, _ = _device_count(bucketed, stream, 8)
        assert lib.fkb_phase_times(ctx, ctypes.byref(ms)) == 0
        assert ms[0] > 0 and ms[1] > 0 and ms[2] > 0  # k = 8: top-bit buckets, so the fold kernels run too
        got11, _ = _device_count(bucketed, stream, 11)
        assert lib.fkb_phase_times(ctx, ctypes.byref(ms)) == 0
        assert ms[0] > 0 and ms[1] > 0 and ms[2] < 0.05  # core buckets: nothing between pass 2 and the end
    finally:
        assert lib.fkb_set_option(ctx, b"phase_events", 0) == 0
    assert np.array_equal(got.table, want.table)


def test_bucketed_and_direct_agree_on_config2(counter, harness):
    from findkmer_b200 import synth
    stream = counter.strip(synth.render(synth.config2()))
    for k in (6, 11):
        counter.set_variant(1)
        a, _ = _device_count(counter, stream, k)
        counter.set_variant(2)
        b, _ = _device_count(counter, stream, k)
        counter.set_variant(0)
        assert np.array_equal(a.table, b.table) and a.raw is not None
        assert (a.n_kmers, a.base_count, a.node_count, a.unknown_chars) == (b.n_kmers, b.base_count, b.node_count, b.unknown_chars)
        assert_counts_equal(b, harness.oracle_count_stream(stream, k))


# ---- the device loader (stream compaction on the GPU; used automatically for pinned inputs) ------------------
@pytest.fixture()
def device_loader(counter):
    counter.set_loader(2)
    yield counter
    counter.set_loader(0)
    counter.set_loader_chunk(0)


def test_device_loader_golden_and_micro_vectors(device_loader, harness, golden, test_txt):
    from findkmer_b200 import FindKmerError, _lib
    for k in (1, 6, 11):
        assert_counts_equal(device_loader.count_fasta(test_txt, k), harness.oracle_count_fasta(test_txt, k))
    for rec in golden["micro"]:
        data = rec["input_latin1"].encode("latin1")
        k = rec["k"]
        if rec["hung"] or data == b"":
            with pytest.raises(FindKmerError) as e:
                device_loader.count_fasta(data, k)
            assert e.value.status == (_lib.FKB_ERR_UNTERMINATED_HEADER if rec["hung"] else _lib.FKB_ERR_EMPTY_INPUT)
            continue
        assert_counts_equal(device_loader.count_fasta(data, k), harness.oracle_count_fasta(data, k))


@pytest.mark.parametrize("chunk", [0, 4096, 1 << 20, (1 << 20) + 4096 + 16, 777_777])
def test_device_loader_chunk_chaining(device_loader, harness, chunk):
    """headers, blank lines, CRLF and junk crossing tile (4 KiB) and chunk edges; then a 0xFF that ends the scan"""
    device_loader.set_loader_chunk(chunk)
    data = bytearray(random_fasta(31, 3_000_000))
    for edge in (4096, 8192, 1 << 20, 2 << 20):
        data[edge - 30:edge + 30] = b">a header line that straddles the edge ACGT >> still header\nACGT"[:60]
    data = bytes(data)
    for k in (3, 11):
        assert_counts_equal(device_loader.count_fasta(data, k), harness.oracle_count_fasta(data, k))
    cut = bytearray(data)
    cut[(1 << 20) + 10] = 0xFF   # inside the straddling header: skipped like any header byte
    cut[2_500_001] = 0xFF        # (almost surely) outside a header: ends the scan
    want = harness.oracle_count_fasta(bytes(cut), 8)
    assert want.bytes_read < len(cut)
    assert_counts_equal(device_loader.count_fasta(bytes(cut), 8), want)


def test_device_loader_equals_host_loader_on_configs(counter, harness):
    from findkmer_b200 import synth
    for lay in (synth.config2(n_bases=5_000_000), synth.config3(n_records=3000), synth.config5(n_bases=4_000_000)):
        data = synth.render(lay)
        counter.set_loader(1)
        a = counter.count_fasta(data, 9)
        counter.set_loader(2)
        b = counter.count_fasta(data, 9)
        counter.set_loader(0)
        assert np.array_equal(a.table, b.table) and a.stream_bytes == b.stream_bytes
        assert (a.n_kmers, a.base_count, a.node_count, a.unknown_chars) == (b.n_kmers, b.base_count, b.node_count, b.unknown_chars)
        assert_counts_equal(b, harness.oracle_count_fasta(data, 9))


def test_device_loader_range_shards(device_loader, harness):
    from findkmer_b200 import synth
    data = synth.render(synth.config5(n_bases=3_000_000))
    want = harness.oracle_count_fasta(data, 8)
    acc = device_loader.new_accumulators(8)
    n = len(data)
    cuts = [0, n // 3 + 17, 2 * n // 3 + 5, n]
    total = 0
    for a, b in zip(cuts[:-1], cuts[1:]):
        lb = max(0, a - 70000)
        sb, stop, eih = device_loader.count_fasta_range(data[lb:b], a - lb, 8, acc)
        assert stop is None
        total += sb
    assert_counts_equal(device_loader.finalize_device(acc, total), want)


def test_multi_k_sweep_in_one_call(counter, harness):
    """the launcher's k = 6..11 sweep (k6thru11fullANDupstream.sh:16-24) as one call: upload + strip once, count per k"""
    from findkmer_b200 import synth
    data = synth.render(synth.config2(n_bases=3_000_000))
    for loader in (1, 2):
        counter.set_loader(loader)
        results = counter.count_fasta_multi(data, range(6, 12))
        counter.set_loader(0)
        for got in results:
            assert_counts_equal(got, harness.oracle_count_fasta(data, got.k))


# ---- loader / generator twins ---------------------------------------------------------------------------
def test_device_generator_is_bit_identical_to_numpy(counter):
    from findkmer_b200 import synth
    for lay in (synth.config2(n_bases=300_007), synth.config3(n_records=300), synth.config5(n_bases=250_000),
                synth.config5(n_bases=250_000).stripped()):
        want = synth.render(lay)
        got = counter.synth_fasta_device(lay).cpu().numpy()
        assert np.array_equal(got, want)
        a, n = 12345, 70001
        assert np.array_equal(counter.synth_fasta_device(lay, a, n).cpu().numpy(), want[a:a + n])


def test_host_pipeline_multi_block_and_eof_alias(counter, harness):
    """> 4 MiB inputs exercise the block pipeline; a 0xFF byte outside a header ends the scan (EOF aliasing)"""
    from findkmer_b200 import synth
    data = synth.render(synth.config2(n_bases=9_000_000)).copy()
    want = harness.oracle_count_fasta(data, 10)
    assert_counts_equal(counter.count_fasta(data, 10), want)
    data[5_000_000] = 0xFF
    want = harness.oracle_count_fasta(data, 10)
    assert want.bytes_read == 5_000_001
    assert_counts_equal(counter.count_fasta(data, 10), want)


def test_host_pipeline_ring_wraps_around(harness):
    """more loader blocks than pinned slots: every slot is reused (a block may only be stripped into a slot whose previous
    copy has been retired), with the count kernels issued per arrived piece; also the option's bounds"""
    from findkmer_b200 import FindKmerError, synth
    from findkmer_b200.engine import KmerCounter
    data = synth.render(synth.config5(n_bases=30_000_000))  # 8 blocks of 4 MiB
    want = harness.oracle_count_fasta(data, 9)
    with KmerCounter(0) as c:
        for bad in (1, 65):
            with pytest.raises(FindKmerError):
                c.set_loader_slots(bad)
        for slots in (2, 3, 64):
            c.set_loader_slots(slots)
            assert_counts_equal(c.count_fasta(data, 9), want)


def test_count_file_and_range_shards(counter, harness, tmp_path):
    from findkmer_b200 import synth
    data = synth.render(synth.config5(n_bases=3_000_000))
    p = tmp_path / "genome.fa"
    data.tofile(p)
    want = harness.oracle_count_fasta(data, 8)
    assert_counts_equal(counter.count_file(str(p), 8), want)
    # three file shards with look-back context, accumulated into one set of device buffers
    acc = counter.new_accumulators(8)
    n = len(data)
    cuts = [0, n // 3 + 17, 2 * n // 3 + 5, n]
    total = 0
    for a, b in zip(cuts[:-1], cuts[1:]):
        lb = max(0, a - 70000)
        sb, stop, eih = counter.count_fasta_range(data[lb:b], a - lb, 8, acc)
        assert stop is None and (not eih or b < n)  # a cut may fall inside a header line; only the file's end must not
        total += sb
    assert_counts_equal(counter.finalize_device(acc, total), want)


# ---- full-size properties (config 4 shape; the oracle cannot run at this size in seconds) -------------
def test_full_size_properties(counter):
    """size-independent checks at 3.1 Gbp, k = 11: conservation (sum of counts == number of windows implied by the
    record layout), exact additivity over a split, and agreement of two different launch decompositions"""
    import torch
    from findkmer_b200 import synth
    n_bases = int(os.environ.get("FKB_TEST_FULL_BASES", 3_100_000_000))
    lay = synth.config4(n_bases=n_bases).stripped()
    k = 11
    d = counter.synth_fasta_device(lay)
    n = d.numel()
    acc = counter.new_accumulators(k)
    counter.count_stream_device(d, k, acc)
    whole = counter.finalize_device(acc, n)
    expect_windows = sum(max(0, lay.record_bases(r) - k + 1) for r in range(lay.n_records))
    assert whole.n_kmers == expect_windows == int(whole.table.sum(dtype=np.uint64))
    assert whole.base_total == n_bases == sum(whole.base_count) and whole.valid_bases == n_bases
    assert whole.node_count == sum(4 ** i for i in range(k + 1))  # every 11-mer occurs ~739 times
    acc2 = counter.new_accumulators(k)
    cut = n // 3 + 1
    counter.count_stream_device(d, k, acc2, 0, cut)
    counter.count_stream_device(d, k, acc2, cut, n)
    split = counter.finalize_device(acc2, n)
    assert np.array_equal(split.table, whole.table)
    assert (split.n_kmers, split.base_count, split.node_count) == (whole.n_kmers, whole.base_count, whole.node_count)
    # the two count paths share no counting code (bucketed: routed 13-mers folded in shared memory; direct: one global red
    # per window, checked against the oracle at small sizes): their full-size tables must be identical bin for bin
    counter.set_variant(1)
    try:
        acc3 = counter.new_accumulators(k)
        counter.count_stream_device(d, k, acc3)
        direct = counter.finalize_device(acc3, n)
    finally:
        counter.set_variant(0)
    assert np.array_equal(direct.table, whole.table)
    assert (direct.n_kmers, direct.base_count, direct.node_count) == (whole.n_kmers, whole.base_count, whole.node_count)
    del d
    torch.cuda.empty_cache()


@pytest.mark.parametrize("name,k", [("config4", 11), ("config5", 8), ("config5", 11), ("config4", 8)])
def test_full_size_against_golden(counter, name, k):
    """BASELINE.json configs 4 and 5 at their FULL size (3.1e9 bases) against the oracle: tests/golden/golden_fullsize.json
    holds sha256(table) and the scalars of a record-parallel run of the C restatement (make_golden_fullsize.py, ~4 min on
    8 cores).  The input is rendered on the device by the bit-identical CUDA twin of synth.py."""
    import hashlib
    import json
    import torch
    from conftest import GOLDEN_DIR
    from findkmer_b200 import synth
    want = json.loads((GOLDEN_DIR / "golden_fullsize.json").read_text())["records"][f"{name}_k{k}"]
    lay = getattr(synth, name)()
    assert (lay.n_bases, lay.n_records, lay.seed) == (want["n_bases"], want["n_records"], want["seed"])
    d = counter.synth_fasta_device(lay.stripped())
    acc = counter.new_accumulators(k)
    counter.count_stream_device(d, k, acc)
    got = counter.finalize_device(acc, d.numel())
    assert got.n_kmers == want["n_kmers"] and got.base_total == want["base_total"]
    assert list(got.base_count) == want["base_count"]
    assert got.node_count == want["node_count"] and got.unknown_chars == want["unknown_chars"]
    assert {str(i): int(got.table[i]) for i in (0, 1, 4 ** k // 3, 4 ** k - 1)} == want["spot"]
    assert hashlib.sha256(np.ascontiguousarray(got.table, dtype="<u4").tobytes()).hexdigest() == want["table_sha256"]
    del d, acc
    torch.cuda.empty_cache()


def test_counter_rollover_is_reported(counter):
    """findKmer.cpp:640-648: a trie node visited 2^32 times ends the reference with COUNTER ROLLOVER DETECTED.  Three poly-A
    records of 1.5e9 bases (each below 2^31: the reference's seqSize is an int) visit the depth-1 node 'A' 4.5e9 times and
    wrap table[AAAAAA]: the library must say FKB_ERR_COUNTER_ROLLOVER through finalize, whichever count path ran."""
    import torch
    from findkmer_b200._lib import FKB_ERR_COUNTER_ROLLOVER, FindKmerError
    n_rec, rec = 3, 1_500_000_000
    d = torch.full((n_rec * (rec + 1),), ord("A"), dtype=torch.uint8, device="cuda:0")
    d[:: rec + 1] = ord(">")
    k = 6
    for variant in (0, 1):
        counter.set_variant(variant)
        try:
            acc = counter.new_accumulators(k)
            counter.count_stream_device(d, k, acc)
            with pytest.raises(FindKmerError) as e:
                counter.finalize_device(acc, d.numel())
            assert e.value.status == FKB_ERR_COUNTER_ROLLOVER
            # the wrapped bin holds the count modulo 2^32, as the reference's unsigned would have
            assert int(acc.table[0].item()) & 0xFFFFFFFF == (n_rec * (rec - k + 1)) % (1 << 32)
        finally:
            counter.set_variant(0)
    # just below the limit nothing is reported: 2 records of 2^31 - 8 bases = 2^32 - 16 visits
    rec2 = (1 << 31) - 8
    d2 = d[: 2 * (rec2 + 1)]
    d2[:] = ord("A")
    d2[:: rec2 + 1] = ord(">")
    d2[5] = ord("C"); d2[6] = ord("G"); d2[7] = ord("T")  # noqa: E702  (all four bases occur)
    acc = counter.new_accumulators(k)
    counter.count_stream_device(d2, k, acc)
    ok = counter.finalize_device(acc, d2.numel())
    assert ok.n_kmers == 2 * (rec2 - k + 1)
    del d, d2, acc
    torch.cuda.empty_cache()


def test_fresh_context_host_paths_after_a_large_device_count(harness):
    """a context whose bucket scratch was sized by a k = 11 device count must not run the k <= 8 bucketed kernels
    without their W-mer table / fold scratch when a HOST entry point is used next (and must size the scratch itself)."""
    import torch
    from findkmer_b200 import synth
    from findkmer_b200.engine import KmerCounter
    with KmerCounter(0) as c:
        lay = synth.config4(n_bases=80_000_000).stripped()
        d = c.synth_fasta_device(lay)
        acc = c.new_accumulators(11)
        c.count_stream_device(d, 11, acc)
        c.finalize_device(acc, d.numel())
        stream = d.cpu().numpy()
        del d
        for k in (8, 7):
            got = c.count_stream(stream, k)   # 64 MiB chunks: above the k = 8 crossover
            sub = stream[: 6_000_000]
            assert np.array_equal(c.count_stream(sub, k).table, harness.oracle_count_stream(sub, k).table)
            expect = sum(max(0, lay.record_bases(r) - k + 1) for r in range(lay.n_records))
            assert got.n_kmers == expect == int(got.table.sum(dtype=np.uint64))
    torch.cuda.empty_cache()


def test_large_soft_masked_bucketed_equals_direct(counter):
    """1 Gbp of the config-5 shape (N runs + soft-masked runs: a run boundary in most warp iterations, so the general path of
    pass 1 -- masks, warp-cooperative run events, junk-cursor atomics, re-read on leaving a clean stretch -- carries the load):
    the bucketed path and the direct kernel must agree bin for bin and scalar for scalar, for a core-bucket k and a top-bit k"""
    import torch
    from findkmer_b200 import synth
    lay = synth.config5(n_bases=int(os.environ.get("FKB_TEST_MASKED_BASES", 1_000_000_000))).stripped()
    d = counter.synth_fasta_device(lay)
    n = d.numel()
    for k in (11, 8):
        acc = counter.new_accumulators(k)
        counter.count_stream_device(d, k, acc)
        auto = counter.finalize_device(acc, n)
        counter.set_variant(1)
        try:
            acc2 = counter.new_accumulators(k)
            counter.count_stream_device(d, k, acc2)
            direct = counter.finalize_device(acc2, n)
        finally:
            counter.set_variant(0)
        assert np.array_equal(auto.table, direct.table)
        assert (auto.n_kmers, auto.base_total, auto.base_count, auto.node_count, auto.unknown_chars, auto.valid_bases, auto.runs_ge_k) == \
               (direct.n_kmers, direct.base_total, direct.base_count, direct.node_count, direct.unknown_chars, direct.valid_bases, direct.runs_ge_k)
        assert auto.n_kmers == int(auto.table.sum(dtype=np.uint64)) and 0 < auto.unknown_chars < n
        del acc, acc2
    del d
    torch.cuda.empty_cache()


# ---- several GPUs behind the C boundary (skipped on a one-GPU box) --------------------------------------------
def test_multi_gpu_c_abi_and_cli(harness, tmp_path):
    """fkb_count_fasta_host_gpus / `findKmer -g N`: contiguous shards on N GPUs, tables summed on GPU 0 over peer access --
    bit-identical to the oracle (and the CLI's files byte-identical to the reference binary's), cuts falling anywhere (inside
    header lines, inside N runs), a terminating byte 0xFF in the middle shard, every k path."""
    import subprocess
    from conftest import ROOT
    from findkmer_b200 import synth
    from findkmer_b200.engine import count_fasta_gpus, device_count
    n_dev = device_count()
    if n_dev < 2:
        pytest.skip("needs at least two GPUs")
    devices = list(range(min(n_dev, 4)))
    data = synth.render(synth.config5(n_bases=3_000_000))
    for k in (6, 8, 11, 13):
        assert_counts_equal(count_fasta_gpus(data, k, devices), harness.oracle_count_fasta(data, k))
    assert_counts_equal(count_fasta_gpus(data, 11, devices[:2]), harness.oracle_count_fasta(data, 11))
    big = synth.render(synth.config4(n_bases=60_000_000))  # large enough for the bucketed kernels on every shard
    assert_counts_equal(count_fasta_gpus(big, 11, devices[:2]), harness.oracle_count_fasta(big, 11))
    records = synth.render(synth.config3(n_records=3000))  # a header line every kilobase: cuts land inside headers
    assert_counts_equal(count_fasta_gpus(records, 9, devices), harness.oracle_count_fasta(records, 9))
    stopped = data.copy()
    stopped[len(stopped) // 2 + 12345] = 0xFF  # ends the scan in a middle shard: later shards must count nothing
    assert_counts_equal(count_fasta_gpus(stopped, 8, devices), harness.oracle_count_fasta(stopped, 8))
    exe = ROOT / "findkmer_b200" / "bin" / "findKmer"
    if exe.exists() and harness.reference_available():
        ref = harness.run_reference(data, 8, name="genome.fa")
        data.tofile(tmp_path / "genome.fa")
        run = subprocess.run([str(exe), "-q", "1", "-k", "8", "-g", str(len(devices)), "-p", "genome.fa"], cwd=tmp_path, capture_output=True,
                             text=True, timeout=300)
        assert run.returncode == 0, run.stderr[-500:]
        assert (tmp_path / ref.csv_name).read_bytes() == ref.csv
        assert (tmp_path / ref.stats_name).read_bytes() == ref.stats


# ---- the drop-in program against the untouched reference binary ---------------------------------------------
def test_cli_binary_is_a_drop_in(harness, tmp_path):
    """findkmer_b200/bin/findKmer vs oracle/_ref/findKmer: same flags, same output file names, identical bytes"""
    import subprocess
    from conftest import ROOT
    from findkmer_b200 import synth
    exe = ROOT / "findkmer_b200" / "bin" / "findKmer"
    if not exe.exists() or not harness.reference_available():
        pytest.skip("CLI or reference binary not built")
    data = synth.render(synth.config5(n_bases=700_000))
    for k, z, export in ((6, None, None), (8, 2, None), (11, None, "mine.csv")):
        ref = harness.run_reference(data, k, z=z, name="genome.fa", export=export)
        work = tmp_path / f"cli_{k}"
        work.mkdir()
        data.tofile(work / "genome.fa")
        argv = [str(exe), "-q", "1", "-k", str(k), "-p", "genome.fa"]
        if z is not None:
            argv += ["-z", str(z)]
        if export:
            argv += ["-e", export, "-t", "mine.tsv"]
        run = subprocess.run(argv, cwd=work, capture_output=True, text=True, timeout=300)
        assert run.returncode == 0, run.stderr[-500:]
        assert (work / ref.csv_name).read_bytes() == ref.csv
        if export:  # --tsv extension: the CSV's rows, tab-separated with the count in column 2 (mergeFile4GNUPLOT.pl's layout)
            rows = ref.csv.split(b"\n")[1:]
            lines = (work / "mine.tsv").read_bytes().split(b"\n")
            assert len(lines) - 1 == len(rows) and lines[-1] == b""
            f = rows[len(rows) // 2].split(b", ")
            assert lines[len(rows) // 2].split(b"\t") == [f[0], f[1], f[3], f[2]] + f[4:]
        assert (work / ref.stats_name).read_bytes() == ref.stats
        # the stdout lines scripts may grep (the reference prints them too)
        for phrase in ("ATTEMPTING CONFIGURATION", "!!!Find The KMER!!!", "Reading sequence from file", "Statistics of occurrences",
                       "valid bases total INSIDE sequences >= k.", "tree density.", "Now creating histogram.", "histogram creation finished."):
            assert phrase in run.stdout and phrase in ref.stdout, phrase
        line = [ln for ln in ref.stdout.split("\n") if "tree density" in ln][0]
        assert line in run.stdout
    # error behaviour: empty file -> message + exit 1, header-only CSV left behind (as the reference does)
    work = tmp_path / "cli_empty"
    work.mkdir()
    (work / "empty.fa").write_bytes(b"")
    run = subprocess.run([str(exe), "-q", "1", "-k", "7", "-p", "empty.fa"], cwd=work, capture_output=True, text=True, timeout=120)
    assert run.returncode == 1 and "Sequence File Is Empty" in run.stderr
    assert (work / "7mer_Historam_Of_empty.fa.csv").read_bytes() == b"Sequence, Shannon Entropy h, Shannon Entropy H, Frequency, Z score"
    run = subprocess.run([str(exe), "-h"], cwd=work, capture_output=True, text=True, timeout=60)
    assert run.returncode == 1 and "Usage: findKmer [options]" in run.stdout


def test_cli_k_sweep_extension(harness, tmp_path):
    """`findKmer -K 6-8` == three separate reference runs (k6thru11fullANDupstream.sh's loop in one process)"""
    import subprocess
    from conftest import ROOT
    from findkmer_b200 import synth
    exe = ROOT / "findkmer_b200" / "bin" / "findKmer"
    if not exe.exists() or not harness.reference_available():
        pytest.skip("CLI or reference binary not built")
    data = synth.render(synth.config3(n_records=300))
    data.tofile(tmp_path / "up.fas")
    run = subprocess.run([str(exe), "-q", "1", "-K", "6-8", "-z", "2", "-p", "up.fas"], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert run.returncode == 0, run.stderr[-500:]
    for k in (6, 7, 8):
        ref = harness.run_reference(data, k, z=2, name="up.fas")
        assert (tmp_path / ref.csv_name).read_bytes() == ref.csv
        assert (tmp_path / ref.stats_name).read_bytes() == ref.stats
