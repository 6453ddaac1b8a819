"""The oracle (oracle/kmer_oracle.c) pinned against the reference's own outputs (CPU only)."""
import hashlib

import numpy as np
import pytest

from conftest import random_fasta

# values the untouched reference binary produced on findKmer/test.txt
SURVEY_TABLE = {
    # k: (N, distinct, base counts, base total, trie nodes, sha256 of the CSV)
    6: (2988, 1882, (737, 816, 718, 732), 3003, 3095, "5a4a06af16c737ecbb0227f0dffdafea13683aeca99ad53b5341fb400534567b"),
    7: (2985, 2558, (737, 816, 718, 732), 3003, 5653, "05a1d736db448513751b36885ae1f36e9e72a89eef45f6443ffbeb5e0356b1f8"),
    8: (2982, 2819, (737, 816, 718, 732), 3003, 8468, "6a33eb327ab8386c77ad1db144f9ca8eaee820e81b805df9895331058356a22f"),
    9: (2979, 2902, (737, 816, 718, 732), 3003, 11364, "b83eb4689cda29e87079ab6f285d4682bf1b7c8e0c47ced6572f8f8ee44d8f48"),
    10: (2976, 2932, (737, 816, 718, 732), 3003, 14285, "2906857dfb274c0283d1c5ffee841eb33746a8a0670c2c6384cad11429fef0b7"),
    11: (2973, 2944, (737, 816, 718, 732), 3003, 17215, "c0fbd112cc7ed88d461ab2bce5f946398dbefceb5ae5c9e5ab2274321aeadbfe"),
    1: (3007, 4, (739, 816, 718, 734), 3007, 5, "abf8f0528d6120fa552b0f854da67b97e123bbcd7938d1bd578c329f833f000e"),
    3: (2999, 64, (739, 816, 718, 734), 3007, 85, "b3c13172336add11cea1a204262f4c5af26854c7616d79611f586d6c8d8707f6"),
}


def test_fixture_is_the_reference_fixture(test_txt):
    assert hashlib.sha256(test_txt).hexdigest() == "359e996b0d6db4dff97f3fdc91d4d154ad02491d40690ca663562515c7f5d823"


@pytest.mark.parametrize("k", sorted(SURVEY_TABLE))
def test_oracle_matches_survey_table(harness, test_txt, golden, k):
    n, distinct, bc, bt, nodes, csv_sha = SURVEY_TABLE[k]
    o = harness.oracle_count_fasta(test_txt, k)
    assert o.rc == harness.FKO_OK
    assert (o.n_kmers, int(np.count_nonzero(o.table)), o.base_count, o.base_total, o.node_count) == (n, distinct, bc, bt, nodes)
    # and the committed golden record says the same about the reference run
    g = golden["test_txt"][str(k)]
    assert g["csv_sha256"] == csv_sha
    assert (g["n_kmers"], g["distinct"], tuple(g["base_count"]), g["base_total"], g["node_count"]) == (n, distinct, bc, bt, nodes)


def test_oracle_matches_golden_csv_tables(harness, test_txt):
    from conftest import GOLDEN_DIR
    for k in (6, 11):
        csv = (GOLDEN_DIR / f"{k}mer_Historam_Of_test.txt.csv").read_bytes()
        assert np.array_equal(harness.csv_to_table(csv, k), harness.oracle_count_fasta(test_txt, k).table)


def test_oracle_k16_and_k20_records(harness, test_txt, golden):
    o = harness.oracle_count_fasta(test_txt, 16)
    g = golden["test_txt"]["16"]
    assert (o.n_kmers, o.node_count, list(o.base_count)) == (g["n_kmers"], g["node_count"], g["base_count"])
    assert harness.oracle_count_fasta(test_txt, 20).rc == harness.FKO_ERR_BAD_K  # dense oracle stops at 16


def test_oracle_micro_vectors(harness, golden):
    """Known-answer vectors of the reset rules, produced by the reference binary."""
    for rec in golden["micro"]:
        data = rec["input_latin1"].encode("latin1")
        o = harness.oracle_count_fasta(data, rec["k"])
        if rec["hung"]:
            assert o.rc == harness.FKO_ERR_UNTERMINATED_HDR, rec
        elif data == b"":
            assert o.rc == harness.FKO_ERR_EMPTY, rec
        else:
            assert o.rc == harness.FKO_OK, rec
            assert o.node_count == rec["node_count"], rec
            if rec["ok"]:
                got = [[int(i), int(o.table[i])] for i in np.flatnonzero(o.table)]
                assert got == rec["table_nonzero"], rec
                assert o.base_total == rec["base_total"], rec
                assert list(o.base_count) == rec["base_count"], rec
            else:  # the reference exited in statistics(): some base has zero probability
                assert rec["division_overflow"] and min(o.base_count) == 0 and o.base_total > 0


def test_oracle_random_golden(harness, golden):
    for rec in golden["random"]:
        data = random_fasta(rec["seed"], rec["n"])
        assert hashlib.sha256(data).hexdigest() == rec["input_sha256"]
        o = harness.oracle_count_fasta(data, rec["k"])
        assert o.rc == harness.FKO_OK
        assert rec["ok"]
        assert hashlib.sha256(o.table.tobytes()).hexdigest() == rec["table_sha256"], rec
        assert (o.n_kmers, o.node_count, list(o.base_count), o.base_total) == (rec["n_kmers"], rec["node_count"], rec["base_count"], rec["base_total"])


def test_oracle_against_reference_binary_live(harness):
    """Where oracle/_ref/findKmer exists (here and, prebuilt, on the GPU box): fresh seeded inputs."""
    if not harness.reference_available():
        pytest.skip("reference binary not built")
    for seed in (101, 102):
        data = random_fasta(seed, 20000)
        for k in (4, 9, 13):
            r = harness.run_reference(data, k)
            o = harness.oracle_count_fasta(data, k)
            assert r.ok and o.rc == 0
            assert np.array_equal(harness.csv_to_table(r.csv, k), o.table)
            assert (r.node_count, r.base_count, r.base_total) == (o.node_count, o.base_count, o.base_total)


def test_oracle_stream_form_equals_fasta_form(harness):
    data = random_fasta(7, 30000)
    rc, stream = harness.oracle_strip(data)
    assert rc == 0 and b"\n" not in bytes(stream)
    for k in (3, 11):
        a, b = harness.oracle_count_fasta(data, k), harness.oracle_count_stream(stream, k)
        assert np.array_equal(a.table, b.table)
        assert (a.n_kmers, a.base_count, a.base_total, a.node_count, a.unknown_chars) == (b.n_kmers, b.base_count, b.base_total, b.node_count, b.unknown_chars)
