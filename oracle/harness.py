"""oracle/harness.py -- TEST INFRASTRUCTURE, NOT PRODUCT.

Python access to the two parity checkers:

* ``oracle_count_fasta`` / ``oracle_count_stream`` / ``oracle_strip`` -- the C restatement
  (``oracle/kmer_oracle.c``, built to ``oracle/_ref/liboracle.so``).
* ``run_reference`` -- the UNTOUCHED reference binary (``oracle/_ref/findKmer`` /
  ``findKmer_probe``, compiled by ``oracle/Makefile`` from /root/reference) run the only way it
  works: inside the input's directory, bare file name, ``-q 1``, exit status
  139/134 ignored, completion judged by the "histogram creation finished." stdout line.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  Nothing here reads /root/reference at run time.
"""
from __future__ import annotations

import ctypes
import os
import re
import shutil
import subprocess
import tempfile
import time
from dataclasses import dataclass, field
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REF_DIR = HERE / "_ref"
LIB_PATH = REF_DIR / "liboracle.so"
REF_BIN = REF_DIR / "findKmer"
REF_PROBE = REF_DIR / "findKmer_probe"

FKO_OK, FKO_ERR_EMPTY, FKO_ERR_UNTERMINATED_HDR, FKO_ERR_ROLLOVER, FKO_ERR_BAD_K, FKO_ERR_NOMEM = range(6)
FKO_MAX_K = 16


class _Result(ctypes.Structure):
    _fields_ = [
        ("n_kmers", ctypes.c_uint64),
        ("base_total", ctypes.c_uint64),
        ("base_count", ctypes.c_uint32 * 4),
        ("node_count", ctypes.c_uint64),
        ("unknown_chars", ctypes.c_uint64),
        ("bytes_read", ctypes.c_uint64),
    ]


@dataclass
class OracleResult:
    rc: int
    table: np.ndarray
    n_kmers: int = 0
    base_total: int = 0
    base_count: tuple = (0, 0, 0, 0)
    node_count: int = 0
    unknown_chars: int = 0
    bytes_read: int = 0


_lib = None


def build(force: bool = False) -> None:
    """Compile the checkers (liboracle.so always; the reference binary when its sources exist)."""
    if force or not LIB_PATH.exists() or LIB_PATH.stat().st_mtime < (HERE / "kmer_oracle.c").stat().st_mtime:
        subprocess.run(["make", "-C", str(HERE), "-s", "_ref/liboracle.so"], check=True)
    if Path("/root/reference/findKmer/src/findKmer.cpp").exists():
        subprocess.run(["make", "-C", str(HERE), "-s", "ref"], check=True)


def lib():
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            build()
        _lib = ctypes.CDLL(str(LIB_PATH))
        u8p = ctypes.c_void_p
        _lib.fko_count_fasta.argtypes = [u8p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p, ctypes.POINTER(_Result)]
        _lib.fko_count_fasta.restype = ctypes.c_int
        _lib.fko_count_stream.argtypes = _lib.fko_count_fasta.argtypes
        _lib.fko_count_stream.restype = ctypes.c_int
        _lib.fko_strip.argtypes = [u8p, ctypes.c_size_t, u8p, ctypes.POINTER(ctypes.c_size_t)]
        _lib.fko_strip.restype = ctypes.c_int
    return _lib


def _as_u8(data) -> np.ndarray:
    if isinstance(data, (bytes, bytearray, memoryview)):
        return np.frombuffer(bytes(data), dtype=np.uint8)
    arr = np.ascontiguousarray(data)
    assert arr.dtype == np.uint8
    return arr


def _count(fn, data, k: int) -> OracleResult:
    arr = _as_u8(data)
    if not (1 <= k <= FKO_MAX_K):
        return OracleResult(FKO_ERR_BAD_K, np.zeros(0, np.uint32))
    table = np.zeros(4 ** k, dtype=np.uint32)
    res = _Result()
    rc = fn(arr.ctypes.data if arr.size else None, arr.size, k, table.ctypes.data, ctypes.byref(res))
    return OracleResult(rc, table, res.n_kmers, res.base_total, tuple(res.base_count), res.node_count,
                        res.unknown_chars, res.bytes_read)


def oracle_count_fasta(data, k: int) -> OracleResult:
    """Count a raw sequence/FASTA byte buffer exactly as the reference scans a file."""
    return _count(lib().fko_count_fasta, data, k)


def oracle_count_stream(data, k: int) -> OracleResult:
    """Count an already-stripped stream (no newlines; headers are lone '>' bytes)."""
    return _count(lib().fko_count_stream, data, k)


def oracle_strip(data):
    """Stream contract on the CPU: returns (rc, stripped bytes as np.uint8)."""
    arr = _as_u8(data)
    out = np.empty(max(arr.size, 1), dtype=np.uint8)
    n = ctypes.c_size_t(0)
    rc = lib().fko_strip(arr.ctypes.data if arr.size else None, arr.size, out.ctypes.data, ctypes.byref(n))
    return rc, out[: n.value].copy()


# ------------------------------------------------------------------------------------------------
# the untouched reference binary
# ------------------------------------------------------------------------------------------------
_LETTER = {ord("A"): 0, ord("C"): 1, ord("G"): 2, ord("T"): 3}


def kmer_to_code(s: str) -> int:
    c = 0
    for ch in s:
        c = (c << 2) | "ACGT".index(ch)
    return c


def code_to_kmer(code: int, k: int) -> str:
    return "".join("ACGT"[(code >> (2 * (k - 1 - i))) & 3] for i in range(k))


def parse_csv_rows(csv: bytes):
    """Split a findKmer CSV into its rows [(kmer, h, H, freq, z-or-None)] (header dropped)."""
    lines = csv.decode("latin1").split("\n")
    rows = []
    for ln in lines[1:]:
        parts = ln.split(", ")
        rows.append((parts[0], parts[1], parts[2], int(parts[3]), parts[4] if len(parts) > 4 else None))
    return lines[0], rows


def csv_to_table(csv: bytes, k: int) -> np.ndarray:
    """Dense 4^k uint32 table from CSV column 4 (zeros <=> absent rows).  Only meaningful without -z."""
    table = np.zeros(4 ** k, dtype=np.uint32)
    data = np.frombuffer(csv, dtype=np.uint8)
    nl = np.flatnonzero(data == 10)
    if nl.size == 0:
        return table
    # vectorised parse: every row starts right after a '\n' with k letters
    starts = nl + 1
    codes = np.zeros(starts.size, dtype=np.int64)
    lut = np.full(256, -1, dtype=np.int64)
    for ch, v in _LETTER.items():
        lut[ch] = v
    for i in range(k):
        codes = (codes << 2) | lut[data[starts + i]]
    # frequency = 4th comma-separated field
    text = csv.decode("latin1")
    freqs = np.fromiter((int(ln.split(", ")[3]) for ln in text.split("\n")[1:]), dtype=np.int64, count=starts.size)
    table[codes] = freqs.astype(np.uint32)
    return table


@dataclass
class RefRun:
    ok: bool                 # reached "histogram creation finished."
    exit_code: int
    stdout: str
    csv: bytes = b""
    stats: bytes = b""
    csv_name: str = ""
    stats_name: str = ""
    node_count: int | None = None
    base_count: tuple | None = None
    base_total: int | None = None
    count_seconds: float | None = None   # between "Reading sequence from file" and "Statistics of occurrences"
    total_seconds: float | None = None
    hung: bool = False
    extra: dict = field(default_factory=dict)


def reference_available() -> bool:
    return REF_BIN.exists() and os.access(REF_BIN, os.X_OK)


def run_reference(data, k: int | None = None, z: int | None = None, *, name: str = "in.fa", probe: bool = True,
                  export: str | None = None, timeout: float = 600.0, time_phases: bool = False,
                  keep_dir: str | None = None, input_path: str | None = None) -> RefRun:
    """Run the untouched reference on `data` (bytes / uint8 array) -- or on an existing file
    `input_path` (hard-linked/copied into the scratch directory) -- and collect its outputs."""
    binary = REF_PROBE if (probe and REF_PROBE.exists()) else REF_BIN
    if not binary.exists():
        raise FileNotFoundError(f"{binary} missing: run `make -C oracle` where /root/reference exists")
    work = Path(keep_dir) if keep_dir else Path(tempfile.mkdtemp(prefix="findkmer_ref_"))
    work.mkdir(parents=True, exist_ok=True)
    try:
        target = work / name
        if input_path is not None:
            try:
                os.link(input_path, target)
            except OSError:
                shutil.copyfile(input_path, target)
        else:
            arr = _as_u8(data)
            arr.tofile(target)
        argv = [str(binary), "-q", "1"]
        if k is not None:
            argv += ["-k", str(k)]
        if z is not None:
            argv += ["-z", str(z)]
        if export is not None:
            argv += ["-e", export]
        argv += ["-p", name]
        env = dict(os.environ, FINDKMER_PROBE_OUT="probe.txt")
        t0 = time.perf_counter()
        hung = False
        count_seconds = None
        if time_phases:
            # A pty makes the child's stdout line-buffered, so the two phase markers
            # ("Reading sequence from file" ... "Statistics of occurrences", main() :1310 and
            # statistics() :512) can be time-stamped as they appear.
            import pty
            import select
            master, slave = pty.openpty()
            proc = subprocess.Popen(argv, cwd=work, env=env, stdout=slave, stderr=subprocess.DEVNULL,
                                    stdin=subprocess.DEVNULL)
            os.close(slave)
            chunks = []
            seen = b""
            t_read = t_stats = None
            deadline = t0 + timeout
            while True:
                if time.perf_counter() > deadline:
                    proc.kill()
                    hung = True
                    break
                r, _, _ = select.select([master], [], [], 0.05)
                if r:
                    try:
                        b = os.read(master, 65536)
                    except OSError:
                        b = b""
                    if not b:
                        break
                    now = time.perf_counter()
                    chunks.append(b)
                    seen = (seen + b)[-4096:] if t_stats else seen + b
                    if t_read is None and b"Reading sequence from file" in seen:
                        t_read = now
                    if t_stats is None and b"Statistics of occurrences" in seen:
                        t_stats = now
                elif proc.poll() is not None:
                    break
            proc.wait()
            os.close(master)
            out_lines = b"".join(chunks).decode("latin1").replace("\r\n", "\n")
            if t_read is not None and t_stats is not None:
                count_seconds = t_stats - t_read
        else:
            proc = subprocess.Popen(argv, cwd=work, env=env, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                                    stdin=subprocess.DEVNULL, text=True, errors="replace")
            try:
                out_lines, _ = proc.communicate(timeout=timeout)
            except subprocess.TimeoutExpired:
                proc.kill()
                out_lines, _ = proc.communicate()
                hung = True
        t1 = time.perf_counter()
        stdout = out_lines
        keff = k if k else 7
        stats_name = f"{keff}mer_Base_Stats_Of_{name}.txt"
        if export is not None:
            csv_name = export
        else:
            csv_name = f"{keff}mer_Historam_Of_{name}{'zScoreFiltered' if z is not None else ''}.csv"
        run = RefRun(ok=("histogram creation finished." in stdout), exit_code=proc.returncode, stdout=stdout,
                     csv_name=csv_name, stats_name=stats_name, hung=hung, total_seconds=t1 - t0,
                     count_seconds=count_seconds)
        if (work / csv_name).exists():
            run.csv = (work / csv_name).read_bytes()
        if (work / stats_name).exists():
            run.stats = (work / stats_name).read_bytes()
        if (work / "probe.txt").exists():
            m = re.search(r"nodeCounter=(\d+)", (work / "probe.txt").read_text())
            if m:
                run.node_count = int(m.group(1))
        m = re.search(r"respectively: \n(\d+), .*\n(\d+), .*\n(\d+), .*\n(\d+), .*\nFound (\d+) valid bases", stdout)
        if m:
            run.base_count = tuple(int(m.group(i)) for i in range(1, 5))
            run.base_total = int(m.group(5))
        return run
    finally:
        if not keep_dir:
            shutil.rmtree(work, ignore_errors=True)
