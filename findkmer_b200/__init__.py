"""findkmer_b200 -- B200-native (sm_100a) drop-in for the counting path of soundude462/findKmer.

Only what the hot path needs lives here: ``csrc/`` (CUDA kernels, the C ABI, the C++ host loader,
writer and command line program), the ctypes binding, the Python mirror of the seam, the sharded
(multi-GPU) driver and the synthetic input generators.
"""
from ._lib import FindKmerError, FKB_MAX_K  # noqa: F401

__all__ = ["FindKmerError", "FKB_MAX_K"]
