"""Build the native pieces in-tree with explicit nvcc / g++ commands (sm_100a only).

    python -m findkmer_b200.build            # library + CLI + micro-benchmarks

Outputs (git-ignored, but they travel to the GPU box with the snapshot):
    findkmer_b200/libfindkmer_b200.so   the extern "C" library of include/findkmer_b200.h
    findkmer_b200/bin/findKmer          the drop-in command line program
    findkmer_b200/bin/fkb_ubench        atomic-throughput micro-benchmarks (profiling aid)
    findkmer_b200/bin/fkb_ubench_bulk   staging-row flush: LSU copies vs cp.async.bulk (profiling aid)
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
OBJ = ROOT / "build" / "obj"
LIB = PKG / "libfindkmer_b200.so"
BIN = PKG / "bin"

NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ARCH + ["-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-I", str(ROOT / "include"), "-I", str(CSRC)]
# the writer must not be compiled with fast-math or FMA contraction: its output is compared byte for byte
CXX_FLAGS = ["-O2", "-fPIC", "-std=c++17", "-Wall", "-Wextra", "-ffp-contract=off", "-I", str(ROOT / "include"), "-I", str(CSRC)]

CU_SOURCES = ["fkb_kernels.cu", "fkb_bucket.cu", "fkb_bucket2.cu", "fkb_smallk.cu", "fkb_strip.cu", "fkb_api.cu"]
CXX_SOURCES = ["fkb_loader.cpp", "fkb_writer.cpp"]


def _stale(target: Path, deps) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(d).stat().st_mtime > t for d in deps)


def _run(cmd, verbose):
    if verbose:
        print(" ".join(str(c) for c in cmd), flush=True)
    subprocess.run([str(c) for c in cmd], check=True)


def build(verbose: bool = False, force: bool = False) -> Path:
    OBJ.mkdir(parents=True, exist_ok=True)
    BIN.mkdir(parents=True, exist_ok=True)
    headers = list(CSRC.glob("*.h")) + list(CSRC.glob("*.cuh")) + [ROOT / "include" / "findkmer_b200.h", Path(__file__)]
    objs = []
    for src in CU_SOURCES:
        o = OBJ / (src + ".o")
        if force or _stale(o, [CSRC / src] + headers):
            _run([NVCC] + NVCC_FLAGS + ["-c", CSRC / src, "-o", o], verbose)
        objs.append(o)
    for src in CXX_SOURCES:
        o = OBJ / (src + ".o")
        if force or _stale(o, [CSRC / src] + headers):
            _run(["g++"] + CXX_FLAGS + ["-c", CSRC / src, "-o", o], verbose)
        objs.append(o)
    if force or _stale(LIB, objs):
        _run([NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-lpthread"], verbose)
    main_src = CSRC / "findkmer_main.cpp"
    exe = BIN / "findKmer"
    if main_src.exists() and (force or _stale(exe, [main_src, LIB] + headers)):
        _run(["g++"] + CXX_FLAGS + [main_src, "-o", exe, "-L", PKG, "-lfindkmer_b200", "-Wl,-rpath,$ORIGIN/..", "-lpthread"], verbose)
    for name in ("fkb_ubench", "fkb_ubench_bulk", "fkb_ubench_stage"):
        ub_src = CSRC / (name + ".cu")
        ub = BIN / name
        if ub_src.exists() and (force or _stale(ub, [ub_src] + headers)):
            _run([NVCC] + NVCC_FLAGS + [ub_src, "-o", ub], verbose)
    return LIB


def build_variant(name: str, defines, verbose: bool = False, only=None) -> Path:
    """A tuning variant of the library (profiling only): same sources, extra -D flags, written to build/variants/.
    `only`: the sources the flags matter for (the others are taken from the regular build)."""
    vdir = ROOT / "build" / "variants" / name
    vdir.mkdir(parents=True, exist_ok=True)
    objs = []
    for src in CU_SOURCES:
        if only is not None and src not in only:
            objs.append(OBJ / (src + ".o"))
            continue
        o = vdir / (src + ".o")
        _run([NVCC] + NVCC_FLAGS + [f"-D{d}" for d in defines] + ["-c", CSRC / src, "-o", o], verbose)
        objs.append(o)
    for src in CXX_SOURCES:
        objs.append(OBJ / (src + ".o"))
    out = PKG / f"libfindkmer_b200_{name}.so"
    _run([NVCC] + ARCH + ["-shared", "-o", out] + objs + ["-lpthread"], verbose)
    return out


if __name__ == "__main__":
    build(verbose=True, force="--force" in sys.argv)
    print(LIB)
