"""Host-side Python mirror of the counting path, over the C ABI (include/findkmer_b200.h).

The product's host code is C++ (csrc/findkmer_main.cpp is the drop-in ``findKmer`` program); this
module is the same seam for Python callers, tests and bench.py.  It mirrors the hand-over of the
reference's ``findKmer()`` (findKmer/src/findKmer.cpp:962-1069): a count table plus
``baseCounter`` / ``baseStatistics[4].Count`` / ``TotalNumSequencesN`` / ``nodeCounter``.

PyTorch is plumbing only (device memory, streams, torch.distributed); every number is produced by
the CUDA kernels behind the C ABI.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import _lib
from ._lib import FkbCounts, FkbPartials, FindKmerError


@dataclass
class KmerCounts:
    """What findKmer() hands to statistics()/histo_recursive(), on the dense table."""
    k: int
    table: np.ndarray            # uint32[4^k]; index = sum b_i 4^(k-1-i), A0 C1 G2 T3
    n_kmers: int                 # TotalNumSequencesN
    base_total: int              # baseCounter
    base_count: tuple            # baseStatistics[i].Count
    node_count: int              # nodeCounter
    unknown_chars: int
    stream_bytes: int
    runs_ge_k: int
    valid_bases: int
    raw: Optional[FkbCounts] = None

    @staticmethod
    def from_struct(k: int, table, c: FkbCounts) -> "KmerCounts":
        copy = FkbCounts.from_buffer_copy(bytes(c))
        return KmerCounts(k, table, c.n_kmers, c.base_total, tuple(c.base_count), c.node_count, c.unknown_chars,
                          c.stream_bytes, c.runs_ge_k, c.valid_bases, copy)


def _host_view(data):
    """(address, nbytes, keep-alive) of bytes / numpy uint8 / CPU torch uint8 without copying when possible."""
    if isinstance(data, (bytes, bytearray)):
        arr = np.frombuffer(data, dtype=np.uint8)
        return (arr.ctypes.data if arr.size else None), arr.size, (data, arr)
    if isinstance(data, np.ndarray):
        arr = np.ascontiguousarray(data)
        if arr.dtype != np.uint8:
            raise TypeError("sequence data must be uint8")
        return (arr.ctypes.data if arr.size else None), arr.size, arr
    try:
        import torch
        if isinstance(data, torch.Tensor):
            if data.device.type != "cpu" or data.dtype != torch.uint8 or not data.is_contiguous():
                raise TypeError("host sequence tensors must be contiguous CPU uint8")
            return (data.data_ptr() if data.numel() else None), data.numel(), data
    except ImportError:  # pragma: no cover
        pass
    raise TypeError(f"unsupported sequence container {type(data)!r}")


class DeviceAccumulators:
    """Per-GPU accumulators of one counting job: table, short-run prefix flags, scalar partials.
    Table and flags share ONE int32 buffer so that a multi-GPU run combines them with a single sum-reduce
    (a flag byte is 0/1 per rank, so byte-wise sums cannot carry for up to 255 ranks and "present" == non-zero);
    the 12 partial scalars are a second, tiny sum-reduce."""

    def __init__(self, k: int, device):
        import torch
        lib = _lib.load()
        self.k = k
        n_table, n_flags = lib.fkb_table_entries(k), lib.fkb_prefix_flags_bytes(k)
        self.buf = torch.zeros(n_table + (n_flags + 3) // 4, dtype=torch.int32, device=device)
        self.table = self.buf[:n_table]                                   # uint32 bit patterns
        self.flags = self.buf[n_table:].view(torch.uint8)[:n_flags]
        self.partials = torch.zeros(ctypes.sizeof(FkbPartials) // 8, dtype=torch.int64, device=device)

    def zero_(self):
        self.buf.zero_()
        self.partials.zero_()


class KmerCounter:
    """One counting context bound to one GPU (one host thread per context)."""

    def __init__(self, device: int = 0):
        self._lib = _lib.load()
        self._ctx = ctypes.c_void_p()
        status = self._lib.fkb_create(int(device), ctypes.byref(self._ctx))
        if status != _lib.FKB_OK:
            self._ctx = None
            raise FindKmerError(status, "fkb_create failed: no usable sm_100 GPU (this engine has no CPU fallback)")
        self.device = int(device)

    # -- lifetime -------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_ctx", None):
            self._lib.fkb_destroy(self._ctx)
            self._ctx = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, status: int):
        if status != _lib.FKB_OK:
            msg = self._lib.fkb_last_error(self._ctx).decode() or _lib.status_string(status)
            raise FindKmerError(status, msg)

    @property
    def launches(self) -> int:
        return int(self._lib.fkb_launch_count(self._ctx))

    def set_variant(self, variant: int) -> None:
        """0 = automatic, 1 = direct kernel only, 2 = bucketed kernels whenever k and the range allow."""
        self._check(self._lib.fkb_set_option(self._ctx, b"variant", int(variant)))

    def set_loader(self, mode: int) -> None:
        """0 = automatic (pinned input -> device strip, pageable -> host strip threads), 1 = host loader, 2 = device loader."""
        self._check(self._lib.fkb_set_option(self._ctx, b"loader", int(mode)))

    def set_loader_slots(self, n_slots: int) -> None:
        """size of the host loader's ring of pinned 4 MiB slots (2..64; default 24)"""
        self._check(self._lib.fkb_set_option(self._ctx, b"loader_slots", int(n_slots)))

    def set_loader_chunk(self, nbytes: int) -> None:
        self._check(self._lib.fkb_set_option(self._ctx, b"loader_chunk", int(nbytes)))

    def device_info(self) -> dict:
        sm, ma, mi, hbm = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_size_t()
        self._check(self._lib.fkb_device_info(self._ctx, ctypes.byref(sm), ctypes.byref(ma), ctypes.byref(mi), ctypes.byref(hbm)))
        return {"sm_count": sm.value, "cc": (ma.value, mi.value), "hbm_bytes": hbm.value}

    # -- host loader ----------------------------------------------------------------------------
    @staticmethod
    def strip(data, n_threads: int = 0) -> np.ndarray:
        """The stream contract on the host (needs no GPU): returns the stripped stream."""
        lib = _lib.load()
        addr, n, keep = _host_view(data)
        out = np.empty(max(n, 1), dtype=np.uint8)
        out_len = ctypes.c_size_t(0)
        status = lib.fkb_strip_fasta(addr, n, out.ctypes.data, ctypes.byref(out_len), n_threads)
        if status != _lib.FKB_OK:
            raise FindKmerError(status, _lib.status_string(status))
        return out[: out_len.value]

    # -- host-buffer (end-to-end) paths ------------------------------------------------------------
    def _host_call(self, fn, data, k: int, table: Optional[np.ndarray]) -> KmerCounts:
        if not (1 <= k <= _lib.FKB_MAX_K):
            raise FindKmerError(_lib.FKB_ERR_BAD_K, f"{k} is not a valid value for k")
        addr, n, keep = _host_view(data)
        if table is None:
            table = np.empty(4 ** k, dtype=np.uint32)
        counts = FkbCounts()
        self._check(fn(self._ctx, addr, n, k, table.ctypes.data, ctypes.byref(counts)))
        return KmerCounts.from_struct(k, table, counts)

    def count_fasta(self, data, k: int, table: Optional[np.ndarray] = None) -> KmerCounts:
        """Raw file bytes (exactly what the reference would fgetc()) -> counts.  Host strip + H2D + kernels + D2H."""
        return self._host_call(self._lib.fkb_count_fasta_host, data, k, table)

    def count_stream(self, data, k: int, table: Optional[np.ndarray] = None) -> KmerCounts:
        """An already stripped stream held on the host -> counts."""
        return self._host_call(self._lib.fkb_count_stream_host, data, k, table)

    def count_fasta_multi(self, data, ks) -> list:
        """Several k over one file image: uploaded and stripped once, counted per k (the launcher's k = 6..11 sweep)."""
        ks = [int(k) for k in ks]
        for k in ks:
            if not (1 <= k <= _lib.FKB_MAX_K):
                raise FindKmerError(_lib.FKB_ERR_BAD_K, f"{k} is not a valid value for k")
        addr, n, keep = _host_view(data)
        tables = [np.empty(4 ** k, dtype=np.uint32) for k in ks]
        c_ks = (ctypes.c_int * len(ks))(*ks)
        c_tables = (ctypes.c_void_p * len(ks))(*[t.ctypes.data for t in tables])
        c_counts = (FkbCounts * len(ks))()
        self._check(self._lib.fkb_count_fasta_host_multi(self._ctx, addr, n, c_ks, len(ks), c_tables, c_counts))
        return [KmerCounts.from_struct(k, t, c_counts[i]) for i, (k, t) in enumerate(zip(ks, tables))]

    def count_file(self, path: str, k: int) -> KmerCounts:
        if not (1 <= k <= _lib.FKB_MAX_K):
            raise FindKmerError(_lib.FKB_ERR_BAD_K, f"{k} is not a valid value for k")
        table = np.empty(4 ** k, dtype=np.uint32)
        counts = FkbCounts()
        self._check(self._lib.fkb_count_file(self._ctx, str(path).encode(), k, table.ctypes.data, ctypes.byref(counts)))
        return KmerCounts.from_struct(k, table, counts)

    def pinned_empty(self, nbytes: int) -> np.ndarray:
        """uint8 numpy array over cudaHostAlloc'ed memory (lives as long as the context)."""
        ptr = ctypes.c_void_p()
        self._check(self._lib.fkb_alloc_pinned(self._ctx, nbytes, ctypes.byref(ptr)))
        buf = (ctypes.c_uint8 * max(nbytes, 1)).from_address(ptr.value)
        return np.frombuffer(buf, dtype=np.uint8, count=nbytes)

    # -- device-resident path (the hot path) ----------------------------------------------------------
    def new_accumulators(self, k: int) -> DeviceAccumulators:
        import torch
        return DeviceAccumulators(k, torch.device("cuda", self.device))

    def count_stream_device(self, d_stream, k: int, acc: DeviceAccumulators, begin: int = 0, end: Optional[int] = None) -> None:
        """Accumulate the windows whose last byte lies in [begin, end) of the device-resident stripped stream
        `d_stream` (torch uint8 CUDA tensor) into `acc`, asynchronously on torch's current stream."""
        import torch
        assert d_stream.is_cuda and d_stream.dtype == torch.uint8 and d_stream.is_contiguous()
        end = d_stream.numel() if end is None else end
        st = torch.cuda.current_stream(d_stream.device).cuda_stream
        self._check(self._lib.fkb_count_stream_device(self._ctx, d_stream.data_ptr(), begin, end, k, acc.table.data_ptr(),
                                                      acc.flags.data_ptr(), acc.partials.data_ptr(), st))

    def finalize_device(self, acc: DeviceAccumulators, stream_bytes: int, fetch_table: bool = True) -> KmerCounts:
        """(table, flags, partials) -- after any cross-GPU reduction -- to KmerCounts; synchronises."""
        import torch
        dev = acc.table.device
        d_counts = torch.zeros(ctypes.sizeof(FkbCounts) // 8, dtype=torch.int64, device=dev)
        st = torch.cuda.current_stream(dev).cuda_stream
        self._check(self._lib.fkb_finalize_device(self._ctx, acc.k, acc.table.data_ptr(), acc.flags.data_ptr(), acc.partials.data_ptr(),
                                                  stream_bytes, d_counts.data_ptr(), st))
        host = d_counts.cpu().numpy()
        c = FkbCounts.from_buffer_copy(host.tobytes())
        table = acc.table.cpu().numpy().view(np.uint32) if fetch_table else None
        if c.rollover:
            raise FindKmerError(_lib.FKB_ERR_COUNTER_ROLLOVER, _lib.status_string(_lib.FKB_ERR_COUNTER_ROLLOVER))
        return KmerCounts.from_struct(acc.k, table, c)

    # -- synthetic inputs in HBM --------------------------------------------------------------------
    def synth_fasta_device(self, layout, first_byte: int = 0, n_bytes: Optional[int] = None):
        """Render bytes [first_byte, first_byte+n_bytes) of a findkmer_b200.synth.FastaLayout directly in HBM
        (bit-identical to synth.render).  Returns a torch uint8 CUDA tensor (16-byte aligned storage)."""
        import torch
        offs, base0 = layout.record_offsets()
        total = int(offs[-1])
        n_bytes = total - first_byte if n_bytes is None else n_bytes
        assert 0 <= first_byte and first_byte + n_bytes <= total
        out = torch.empty(n_bytes + 64, dtype=torch.uint8, device=torch.device("cuda", self.device))
        headers = np.frombuffer(b"".join(layout.header(r) for r in range(layout.n_records)), dtype=np.uint8)
        st = torch.cuda.current_stream(out.device).cuda_stream
        self._check(self._lib.fkb_synth_fasta_device(self._ctx, out.data_ptr(), first_byte, n_bytes, layout.n_records, offs.ctypes.data,
                                                     base0.ctypes.data, headers.ctypes.data, layout.header_len, layout.line_width,
                                                     layout.seed & ((1 << 64) - 1), int(layout.n_runs), int(layout.soft_mask), st))
        return out[:n_bytes]

    def count_fasta_range(self, data, own_offset: int, k: int, acc: DeviceAccumulators):
        """One shard of a file (multi-GPU): data = raw bytes with `own_offset` bytes of look-back context in front.
        Accumulates into `acc`; returns (stream_bytes, stop_offset or None, ends_in_header)."""
        addr, n, keep = _host_view(data)
        sb, stop, eih = ctypes.c_uint64(0), ctypes.c_uint64(0), ctypes.c_int(0)
        self._check(self._lib.fkb_count_fasta_host_range(self._ctx, addr, n, own_offset, k, acc.table.data_ptr(), acc.flags.data_ptr(),
                                                         acc.partials.data_ptr(), ctypes.byref(sb), ctypes.byref(stop), ctypes.byref(eih)))
        return sb.value, (None if stop.value == (1 << 64) - 1 else stop.value), bool(eih.value)


def device_count() -> int:
    return int(_lib.load().fkb_device_count())


def count_fasta_gpus(data, k: int, devices) -> KmerCounts:
    """Raw file bytes -> counts on several GPUs of this box (contiguous shards, one host thread and one context per GPU, the
    per-GPU tables summed on devices[0] over NVLink peer access) -- all inside the C library (fkb_count_fasta_host_gpus)."""
    lib = _lib.load()
    if not (1 <= k <= _lib.FKB_MAX_K):
        raise FindKmerError(_lib.FKB_ERR_BAD_K, f"{k} is not a valid value for k")
    addr, n, keep = _host_view(data)
    devs = (ctypes.c_int * len(devices))(*[int(d) for d in devices])
    table = np.empty(4 ** k, dtype=np.uint32)
    counts = FkbCounts()
    err = ctypes.create_string_buffer(512)
    status = lib.fkb_count_fasta_host_gpus(devs, len(devices), addr, n, k, table.ctypes.data, ctypes.byref(counts), err, len(err))
    if status != _lib.FKB_OK:
        raise FindKmerError(status, err.value.decode() or _lib.status_string(status))
    return KmerCounts.from_struct(k, table, counts)


# ------------------------------------------------------------------------------------------------
# writers (host C++ behind the ABI): statistics() and histo_recursive() of the reference
# ------------------------------------------------------------------------------------------------
_libc = None


def _c():
    global _libc
    if _libc is None:
        _libc = ctypes.CDLL(None)
        _libc.fopen.argtypes = [ctypes.c_char_p, ctypes.c_char_p]
        _libc.fopen.restype = ctypes.c_void_p
        _libc.fclose.argtypes = [ctypes.c_void_p]
        _libc.fclose.restype = ctypes.c_int
        _libc.fputs.argtypes = [ctypes.c_char_p, ctypes.c_void_p]
        _libc.fputs.restype = ctypes.c_int
    return _libc


CSV_HEADER = b"Sequence, Shannon Entropy h, Shannon Entropy H, Frequency, Z score"  # findKmer.cpp:79, written without '\n' (:354)


def write_outputs(counts: KmerCounts, csv_path: str, stats_path: str, z_threshold: Optional[int] = None,
                  console_path: Optional[str] = None, n_threads: int = 0, tsv_path: Optional[str] = None) -> int:
    """Write <k>mer_Historam_Of_<file>.csv and <k>mer_Base_Stats_Of_<file>.txt byte-identically to the reference.
    Returns the FKB status (FKB_ERR_ZERO_BASE_PROBABILITY leaves a header-only CSV, as the reference does).
    tsv_path: additionally the tab-separated table mergeFile4GNUPLOT.pl joins (kmer, h, frequency, H[, z])."""
    lib = _lib.load()
    c = _c()
    raw = counts.raw
    csv = c.fopen(str(csv_path).encode(), b"w")
    stats = c.fopen(str(stats_path).encode(), b"w")
    console = c.fopen(str(console_path).encode(), b"w") if console_path else None
    tsv = c.fopen(str(tsv_path).encode(), b"w") if tsv_path else None
    if not csv or not stats or (tsv_path and not tsv):
        raise OSError("cannot open output files")
    try:
        c.fputs(CSV_HEADER, csv)
        probs = (ctypes.c_longdouble * 4)()
        status = lib.fkb_write_base_stats(stats, console, counts.k, ctypes.byref(raw), ctypes.byref(probs))
        if status == _lib.FKB_OK:
            rows = ctypes.c_uint64(0)
            table = np.ascontiguousarray(counts.table, dtype=np.uint32)
            status = lib.fkb_write_histogram(csv, counts.k, table.ctypes.data, ctypes.byref(raw), ctypes.byref(probs),
                                             1 if z_threshold is not None else 0,
                                             ctypes.c_longdouble(float(z_threshold) if z_threshold is not None else 1000.0),
                                             n_threads, ctypes.byref(rows))
            if status == _lib.FKB_OK and tsv:
                status = lib.fkb_write_histogram_tsv(tsv, counts.k, table.ctypes.data, ctypes.byref(raw), ctypes.byref(probs),
                                                     1 if z_threshold is not None else 0,
                                                     ctypes.c_longdouble(float(z_threshold) if z_threshold is not None else 1000.0),
                                                     n_threads, ctypes.byref(rows))
        return status
    finally:
        c.fclose(csv)
        c.fclose(stats)
        if tsv:
            c.fclose(tsv)
        if console:
            c.fclose(console)
