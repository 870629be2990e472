"""TEST INFRASTRUCTURE: loaders for the reference build (oracle/_ref), this repo's C
restatement (oracle/liboracle.so) and the CPU kernel-logic emulation (tests/host_emu)."""
from __future__ import annotations

import ctypes as C
import importlib
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
bwa = importlib.import_module("network-aware-bwa_b200")
abi = bwa.abi

REF_DIR = os.path.join(ROOT, "oracle", "_ref")
REF_BWA = os.path.join(REF_DIR, "bwa")


def have_ref() -> bool:
    return os.path.exists(os.path.join(REF_DIR, "libbwaref.so")) and os.path.exists(
        os.path.join(REF_DIR, "libref_harness.so"))


_ref = None


def ref():
    """(libbwaref, libref_harness) with argtypes set."""
    global _ref
    if _ref is None:
        L = C.CDLL(os.path.join(REF_DIR, "libbwaref.so"), mode=C.RTLD_GLOBAL)
        H = C.CDLL(os.path.join(REF_DIR, "libref_harness.so"))
        PP = C.POINTER(C.POINTER(abi.bwt_t))
        L.bwa_cal_sa_reg_gap.argtypes = [PP, C.c_int, C.POINTER(abi.bwa_seq_t), C.POINTER(abi.gap_opt_t)]
        L.bwa_cal_sa_reg_gap.restype = None
        L.bwt_sa.argtypes = [C.POINTER(abi.bwt_t), C.c_uint32]
        L.bwt_sa.restype = C.c_uint32
        L.bwa_cal_maxdiff.argtypes = [C.c_int, C.c_double, C.c_double]
        L.bwa_cal_maxdiff.restype = C.c_int
        H.refh_aln_batch.argtypes = [PP, C.c_int, C.POINTER(abi.bwa_seq_t), C.POINTER(abi.gap_opt_t), C.c_int]
        H.refh_sa_batch.argtypes = [PP, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        H.refh_sw1.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_int), C.c_void_p]
        H.refh_sw1.restype = C.c_int
        _ref = (L, H)
    return _ref


class RefIndex:
    """bwt_t pair over an index.FMIndex for calling the reference."""

    def __init__(self, idx):
        self.idx = idx
        self.t = [abi.make_bwt_t(idx.bwt[0]), abi.make_bwt_t(idx.bwt[1])]
        self.arr = (C.POINTER(abi.bwt_t) * 2)(C.pointer(self.t[0]), C.pointer(self.t[1]))


def ref_aln(ridx: RefIndex, reads, opt, threads: int = 8):
    """Reference bwa_cal_sa_reg_gap with n_seqs = 1 per read -> (n_aln, max_entries, aln_off, aln)"""
    _, H = ref()
    seqs, keep = abi.make_seqs(reads)
    n = len(seqs)
    H.refh_aln_batch(ridx.arr, n, seqs, C.byref(opt), threads)
    n_aln = np.array([s.n_aln for s in seqs], dtype=np.int32)
    max_entries = np.array([s.max_entries for s in seqs], dtype=np.int32)
    aln_off = np.zeros(n + 1, dtype=np.int64)
    aln_off[1:] = np.cumsum(n_aln)
    aln = np.empty(int(aln_off[n]), dtype=abi.ALN_DTYPE)
    libc = C.CDLL(None)
    libc.free.argtypes = [C.c_void_p]
    for i, s in enumerate(seqs):
        if s.n_aln:
            buf = (C.c_char * (16 * s.n_aln)).from_address(C.addressof(s.aln.contents))
            aln[aln_off[i]:aln_off[i + 1]] = np.frombuffer(buf, dtype=abi.ALN_DTYPE, count=s.n_aln)
        if s.aln:
            libc.free(C.cast(s.aln, C.c_void_p))
    return n_aln, max_entries, aln_off, aln


def ref_sa(ridx: RefIndex, k: np.ndarray, which: np.ndarray, threads: int = 8) -> np.ndarray:
    _, H = ref()
    k = np.ascontiguousarray(k, dtype=np.uint32)
    which = np.ascontiguousarray(which, dtype=np.uint8)
    out = np.empty(k.size, dtype=np.uint32)
    H.refh_sa_batch(ridx.arr, k.size, k.ctypes.data, which.ctypes.data, out.ctypes.data, threads)
    return out


# ---------------------------------------------------------------- kernel-logic emulation
EMU_DIR = os.path.join(ROOT, "tests", "host_emu")
_emu = None


def emu():
    global _emu
    if _emu is None:
        so = os.path.join(EMU_DIR, "libkernel_emu.so")
        srcs = [os.path.join(EMU_DIR, "kernel_emu.cpp"), os.path.join(EMU_DIR, "host_emu_shim.h")] + [
            os.path.join(ROOT, "network-aware-bwa_b200", "csrc", f) for f in ("kernels.cuh", "fmindex.cuh", "hostprep.h")]
        if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
            subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-w", "-I", EMU_DIR, "-o", so, srcs[0]],
                           check=True)
        E = C.CDLL(so)
        E.emu_last_error.restype = C.c_char_p
        E.emu_load_index.argtypes = [C.POINTER(C.POINTER(abi.bwt_t))]
        E.emu_load_index.restype = C.c_void_p
        E.emu_free_index.argtypes = [C.c_void_p]
        E.emu_aln_flat.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(abi.gap_opt_t), C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.c_uint32, C.c_uint32, C.c_int,
                                   C.c_void_p]
        E.emu_free.argtypes = [C.c_void_p]
        E.emu_sa.argtypes = [C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p]
        _emu = E
    return _emu


def emu_aln(h, reads, opt, cap1=1024, aln_cap1=64, n_slots=3):
    E = emu()
    n = reads.n
    bases = np.ascontiguousarray(reads.bases, dtype=np.uint8)
    offs = np.ascontiguousarray(reads.offs, dtype=np.int64)
    n_aln = np.empty(n, dtype=np.int32)
    max_entries = np.empty(n, dtype=np.int32)
    aln_off = np.empty(n + 1, dtype=np.int64)
    pool = C.c_void_p()
    stats = np.zeros(8, dtype=np.uint64)
    rc = E.emu_aln_flat(h, n, bases.ctypes.data, offs.ctypes.data, C.byref(opt), n_aln.ctypes.data,
                        max_entries.ctypes.data, aln_off.ctypes.data, C.byref(pool), cap1, aln_cap1, n_slots,
                        stats.ctypes.data)
    if rc:
        raise RuntimeError(E.emu_last_error().decode())
    tot = int(aln_off[n])
    aln = np.empty(tot, dtype=abi.ALN_DTYPE)
    if tot:
        buf = (C.c_char * (16 * tot)).from_address(pool.value)
        aln[:] = np.frombuffer(buf, dtype=abi.ALN_DTYPE, count=tot)
    E.emu_free(pool)
    return n_aln, max_entries, aln_off, aln, stats


def compare_aln(a, b, label=""):
    """a, b = (n_aln, max_entries, aln_off, aln).  Returns list of mismatch strings."""
    errs = []
    na, ma, oa, aa = a[:4]
    nb, mb, ob, ab = b[:4]
    if not np.array_equal(na, nb):
        bad = np.nonzero(na != nb)[0]
        errs.append(f"{label} n_aln differs for {bad.size} reads, first {bad[:5]}: {na[bad[:5]]} vs {nb[bad[:5]]}")
        return errs
    if not np.array_equal(ma, mb):
        bad = np.nonzero(ma != mb)[0]
        errs.append(f"{label} max_entries differs for {bad.size} reads, first {bad[:5]}: {ma[bad[:5]]} vs {mb[bad[:5]]}")
    if aa.tobytes() != ab.tobytes():
        neq = np.nonzero((aa["info"] != ab["info"]) | (aa["k"] != ab["k"]) | (aa["l"] != ab["l"]) |
                         (aa["score"] != ab["score"]))[0]
        errs.append(f"{label} aln records differ at {neq.size} positions, first {neq[:5]}: {aa[neq[:3]]} vs {ab[neq[:3]]}")
    return errs
