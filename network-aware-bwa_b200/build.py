"""Builds csrc/ into libbwagpu.so IN-TREE with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libbwagpu.so")
SOURCES = ["bwagpu.cu"]
DEPS = ["bwagpu.cu", "kernels.cuh", "search_warp.cuh", "fmindex.cuh", "sw.cuh", "hostprep.h", os.path.join("..", "..", "include", "bwa_gpu.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def build_variant(name: str, defines: list) -> str:
    """Experiment builds (scripts/ab.sh): same sources with -D switches -> variants/libbwagpu_<name>.so,
    selected at run time with BWAGPU_LIB=<path>."""
    out_dir = os.path.join(HERE, "variants")
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, f"libbwagpu_{name}.so")
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-o", out, *[os.path.join(CSRC, s) for s in SOURCES]]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    return out


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile if sources are newer than the library (or no nvcc: use the prebuilt one)."""
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    if not os.path.exists(nvcc):
        if os.path.exists(LIB):
            return LIB
        raise RuntimeError("nvcc not found and no prebuilt libbwagpu.so")
    cmd = [nvcc, *NVCC_FLAGS, "-o", LIB, *[os.path.join(CSRC, s) for s in SOURCES]]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
