"""ctypes binding of include/findkmer_b200.h (the C-ABI shared library built in-tree).

There is no fallback: if the library is missing this module raises, and every counting call fails
loudly when no sm_100 GPU is usable (fkb_create returns FKB_ERR_CUDA).
"""
from __future__ import annotations

import ctypes
from ctypes import POINTER, c_char_p, c_int, c_size_t, c_uint8, c_uint32, c_uint64, c_void_p
from pathlib import Path

PKG = Path(__file__).resolve().parent
LIB_PATH = PKG / "libfindkmer_b200.so"

FKB_MAX_K = 16
FKB_OK = 0
FKB_ERR_EMPTY_INPUT = 1
FKB_ERR_UNTERMINATED_HEADER = 2
FKB_ERR_COUNTER_ROLLOVER = 3
FKB_ERR_BAD_K = 4
FKB_ERR_NOMEM = 5
FKB_ERR_CUDA = 6
FKB_ERR_BAD_ARG = 7
FKB_ERR_IO = 8
FKB_ERR_ZERO_BASE_PROBABILITY = 9

# every symbol include/findkmer_b200.h declares (tests check that the built library exports all of them)
EXPORTS = [
    "fkb_create", "fkb_destroy", "fkb_last_error", "fkb_status_string", "fkb_version", "fkb_device_info", "fkb_set_option", "fkb_phase_times",
    "fkb_strip_fasta", "fkb_alloc_pinned", "fkb_free_pinned",
    "fkb_table_entries", "fkb_prefix_flags_bytes", "fkb_zero_device", "fkb_count_stream_device", "fkb_finalize_device",
    "fkb_count_fasta_host", "fkb_count_stream_host", "fkb_count_fasta_host_range", "fkb_count_file",
    "fkb_count_fasta_host_multi", "fkb_count_file_multi", "fkb_count_fasta_host_gpus", "fkb_count_file_gpus", "fkb_device_count",
    "fkb_write_base_stats", "fkb_write_histogram", "fkb_write_histogram_tsv", "fkb_max_nodes",
    "fkb_synth_fasta_device", "fkb_launch_count",
]


class FkbCounts(ctypes.Structure):
    _fields_ = [
        ("n_kmers", c_uint64),
        ("base_total", c_uint64),
        ("base_count", c_uint64 * 4),
        ("node_count", c_uint64),
        ("unknown_chars", c_uint64),
        ("stream_bytes", c_uint64),
        ("runs_ge_k", c_uint64),
        ("valid_bases", c_uint64),
        ("rollover", c_uint64),
    ]


class FkbPartials(ctypes.Structure):
    _fields_ = [
        ("head_base", c_uint64 * 4),
        ("short_first", c_uint64 * 4),
        ("runs_ge_k", c_uint64),
        ("unknown_chars", c_uint64),
        ("valid_bases", c_uint64),
        ("n_windows", c_uint64),
    ]


class FindKmerError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"findkmer_b200 status {status}: {message}")
        self.status = status


_lib = None


def load() -> ctypes.CDLL:
    """Load the in-tree library (building it first if it has never been built)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        from . import build as _build
        _build.build()
    import os
    lib = ctypes.CDLL(os.environ.get("FKB_LIB", str(LIB_PATH)))  # FKB_LIB: a tuning variant built by build.build_variant()
    ctx = c_void_p
    lib.fkb_create.argtypes = [c_int, POINTER(ctx)]
    lib.fkb_create.restype = c_int
    lib.fkb_destroy.argtypes = [ctx]
    lib.fkb_destroy.restype = None
    lib.fkb_last_error.argtypes = [ctx]
    lib.fkb_last_error.restype = c_char_p
    lib.fkb_status_string.argtypes = [c_int]
    lib.fkb_status_string.restype = c_char_p
    lib.fkb_version.argtypes = []
    lib.fkb_version.restype = c_char_p
    lib.fkb_device_info.argtypes = [ctx, POINTER(c_int), POINTER(c_int), POINTER(c_int), POINTER(c_size_t)]
    lib.fkb_device_info.restype = c_int
    lib.fkb_set_option.argtypes = [ctx, c_char_p, ctypes.c_long]
    lib.fkb_set_option.restype = c_int
    lib.fkb_phase_times.argtypes = [ctx, POINTER(ctypes.c_double * 3)]
    lib.fkb_phase_times.restype = c_int
    lib.fkb_strip_fasta.argtypes = [c_void_p, c_size_t, c_void_p, POINTER(c_size_t), c_int]
    lib.fkb_strip_fasta.restype = c_int
    lib.fkb_alloc_pinned.argtypes = [ctx, c_size_t, POINTER(c_void_p)]
    lib.fkb_alloc_pinned.restype = c_int
    lib.fkb_free_pinned.argtypes = [ctx, c_void_p]
    lib.fkb_free_pinned.restype = c_int
    lib.fkb_table_entries.argtypes = [c_int]
    lib.fkb_table_entries.restype = c_size_t
    lib.fkb_prefix_flags_bytes.argtypes = [c_int]
    lib.fkb_prefix_flags_bytes.restype = c_size_t
    lib.fkb_zero_device.argtypes = [ctx, c_int, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.fkb_zero_device.restype = c_int
    lib.fkb_count_stream_device.argtypes = [ctx, c_void_p, c_uint64, c_uint64, c_int, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.fkb_count_stream_device.restype = c_int
    lib.fkb_finalize_device.argtypes = [ctx, c_int, c_void_p, c_void_p, c_void_p, c_uint64, c_void_p, c_void_p]
    lib.fkb_finalize_device.restype = c_int
    lib.fkb_count_fasta_host.argtypes = [ctx, c_void_p, c_size_t, c_int, c_void_p, POINTER(FkbCounts)]
    lib.fkb_count_fasta_host.restype = c_int
    lib.fkb_count_stream_host.argtypes = [ctx, c_void_p, c_size_t, c_int, c_void_p, POINTER(FkbCounts)]
    lib.fkb_count_stream_host.restype = c_int
    lib.fkb_count_fasta_host_multi.argtypes = [ctx, c_void_p, c_size_t, POINTER(c_int), c_int, POINTER(c_void_p), POINTER(FkbCounts)]
    lib.fkb_count_fasta_host_multi.restype = c_int
    lib.fkb_count_file_multi.argtypes = [ctx, c_char_p, POINTER(c_int), c_int, POINTER(c_void_p), POINTER(FkbCounts)]
    lib.fkb_count_file_multi.restype = c_int
    lib.fkb_count_fasta_host_gpus.argtypes = [POINTER(c_int), c_int, c_void_p, c_size_t, c_int, c_void_p, POINTER(FkbCounts), c_char_p, c_size_t]
    lib.fkb_count_fasta_host_gpus.restype = c_int
    lib.fkb_count_file_gpus.argtypes = [POINTER(c_int), c_int, c_char_p, c_int, c_void_p, POINTER(FkbCounts), c_char_p, c_size_t]
    lib.fkb_count_file_gpus.restype = c_int
    lib.fkb_device_count.argtypes = []
    lib.fkb_device_count.restype = c_int
    lib.fkb_count_file.argtypes = [ctx, c_char_p, c_int, c_void_p, POINTER(FkbCounts)]
    lib.fkb_count_file.restype = c_int
    lib.fkb_write_base_stats.argtypes = [c_void_p, c_void_p, c_int, POINTER(FkbCounts), POINTER(ctypes.c_longdouble * 4)]
    lib.fkb_write_base_stats.restype = c_int
    lib.fkb_write_histogram.argtypes = [c_void_p, c_int, c_void_p, POINTER(FkbCounts), POINTER(ctypes.c_longdouble * 4), c_int,
                                        ctypes.c_longdouble, c_int, POINTER(c_uint64)]
    lib.fkb_write_histogram.restype = c_int
    lib.fkb_write_histogram_tsv.argtypes = lib.fkb_write_histogram.argtypes
    lib.fkb_write_histogram_tsv.restype = c_int
    lib.fkb_max_nodes.argtypes = [c_int]
    lib.fkb_max_nodes.restype = c_uint64
    lib.fkb_synth_fasta_device.argtypes = [ctx, c_void_p, c_uint64, c_uint64, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                           c_uint64, c_int, c_int, c_void_p]
    lib.fkb_count_fasta_host_range.argtypes = [ctx, c_void_p, c_size_t, c_size_t, c_int, c_void_p, c_void_p, c_void_p,
                                               POINTER(c_uint64), POINTER(c_uint64), POINTER(c_int)]
    lib.fkb_count_fasta_host_range.restype = c_int
    lib.fkb_synth_fasta_device.restype = c_int
    lib.fkb_launch_count.argtypes = [ctx]
    lib.fkb_launch_count.restype = c_uint64
    _lib = lib
    return lib


def status_string(status: int) -> str:
    return load().fkb_status_string(status).decode()
