import json
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN_DIR = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden():
    return json.loads((GOLDEN_DIR / "golden.json").read_text())


@pytest.fixture(scope="session")
def test_txt():
    return (GOLDEN_DIR / "test.txt").read_bytes()


@pytest.fixture(scope="session")
def harness():
    """The parity checkers (oracle/ is test infrastructure)."""
    from oracle import harness as H
    H.build()
    return H


@pytest.fixture(scope="session")
def fkb_lib():
    """The built C-ABI library (compiled here with nvcc if missing; loading it needs no GPU)."""
    from findkmer_b200 import _lib, build
    build.build()
    return _lib.load()


@pytest.fixture(scope="session")
def counter(fkb_lib):
    """One counting context on cuda:0 -- fails loudly (no fallback) when there is no B200."""
    from findkmer_b200.engine import KmerCounter
    c = KmerCounter(0)
    yield c
    c.close()


def random_fasta(seed: int, n: int) -> bytes:
    """Same generator as tests/golden/make_golden.py::random_input (junk bytes, headers, CRLF, N, lower case)."""
    rng = np.random.default_rng(seed)
    alphabet = np.frombuffer(b"ACGT" * 12 + b"NNacgt\r \tRY-*.0" + b"\n\n\n", dtype=np.uint8)
    body = alphabet[rng.integers(0, alphabet.size, size=n)].copy()
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


def assert_counts_equal(got, want, check_nodes=True):
    """got: findkmer_b200.engine.KmerCounts, want: oracle OracleResult."""
    assert got.n_kmers == want.n_kmers
    assert got.base_total == want.base_total
    assert tuple(got.base_count) == tuple(want.base_count)
    if check_nodes:
        assert got.node_count == want.node_count
    assert got.unknown_chars == want.unknown_chars
    assert np.array_equal(got.table, want.table)
