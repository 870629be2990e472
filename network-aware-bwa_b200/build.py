"""Builds csrc/ into libbwagpu.so IN-TREE with nvcc for sm_100a (cross-compiles without a GPU).

Each translation unit is compiled to an object of its own (rebuilt only when one of its dependencies changed) and the
objects are linked into the one shared library the C-ABI lives in."""
from __future__ import annotations

import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libbwagpu.so")
HEADER = os.path.join("..", "..", "include", "bwa_gpu.h")
# translation unit -> what it depends on (besides itself)
UNITS = {
    "bwagpu.cu": ["kernels.cuh", "search_warp.cuh", "fmindex.cuh", "sw.cuh", "hostprep.h", HEADER],
    "indexbuild.cu": [HEADER],
    "bgzf.cu": ["bgzf.cuh", HEADER],
}
SOURCES = list(UNITS)
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
]


def _nvcc() -> str:
    return os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


def _mtime(rel: str) -> float:
    return os.path.getmtime(os.path.join(CSRC, rel))


def _obj_of(src: str, tag: str = "") -> str:
    return os.path.join(OBJ, os.path.splitext(src)[0] + tag + ".o")


def _obj_stale(src: str, obj: str) -> bool:
    if not os.path.exists(obj):
        return True
    t = os.path.getmtime(obj)
    return any(_mtime(d) > t for d in [src, *UNITS[src]])


def _compile(src: str, obj: str, defines=(), verbose: bool = False) -> str:
    cmd = [_nvcc(), *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-c", "-o", obj, os.path.join(CSRC, src)]
    if verbose:
        cmd[1:1] = ["-Xptxas", "-v"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    return r.stderr


def _link(objs, out: str) -> None:
    r = subprocess.run([_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", out, *objs], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc link failed:\n" + r.stdout + r.stderr)


def build_variant(name: str, defines: list, units=("bwagpu.cu",)) -> str:
    """Experiment builds (scripts/ab.sh): the named units with -D switches -> variants/libbwagpu_<name>.so,
    selected at run time with BWAGPU_LIB=<path>.  The other units are shared with the default build."""
    out_dir = os.path.join(HERE, "variants")
    os.makedirs(out_dir, exist_ok=True)
    os.makedirs(OBJ, exist_ok=True)
    out = os.path.join(out_dir, f"libbwagpu_{name}.so")
    objs = []
    for src in SOURCES:
        if src in units:
            obj = _obj_of(src, "_" + name)
            _compile(src, obj, defines)
        else:
            obj = _obj_of(src)
            if _obj_stale(src, obj):
                _compile(src, obj)
        objs.append(obj)
    _link(objs, out)
    return out


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile what is stale and link (no nvcc: use the prebuilt library)."""
    if not os.path.exists(_nvcc()):
        if os.path.exists(LIB):
            return LIB
        raise RuntimeError("nvcc not found and no prebuilt libbwagpu.so")
    os.makedirs(OBJ, exist_ok=True)
    todo = [s for s in SOURCES if force or _obj_stale(s, _obj_of(s))]
    if todo:
        with ThreadPoolExecutor(len(todo)) as ex:
            logs = list(ex.map(lambda s: _compile(s, _obj_of(s), (), verbose), todo))
        if verbose:
            print("\n".join(logs))
    objs = [_obj_of(s) for s in SOURCES]
    if todo or not os.path.exists(LIB) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        _link(objs, LIB)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
