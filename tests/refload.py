"""TEST INFRASTRUCTURE: loaders for the reference build (oracle/_ref), this repo's C
restatement (oracle/liboracle.so) and the CPU kernel-logic emulation (tests/host_emu)."""
from __future__ import annotations

import ctypes as C
import importlib
import os
import time
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
    t0 = time.perf_counter()
    H.refh_aln_batch(ridx.arr, n, seqs, C.byref(opt), threads)
    ref_aln.last_batch_s = time.perf_counter() - t0  # the reference call alone (bench.py reports it), without the marshalling around it
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
                                   C.c_void_p, C.c_uint32]
        E.emu_free.argtypes = [C.c_void_p]
        E.emu_sa.argtypes = [C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p]
        _emu = E
    return _emu


def emu_aln(h, reads, opt, cap1=1024, aln_cap1=64, n_slots=3, pool_chunks=64):
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
                        stats.ctypes.data, pool_chunks)
    if rc:
        raise RuntimeError(E.emu_last_error().decode())
    tot = int(aln_off[n])
    aln = np.empty(tot, dtype=abi.ALN_DTYPE)
    if tot:
        buf = (C.c_char * (16 * tot)).from_address(pool.value)
        aln[:] = np.frombuffer(buf, dtype=abi.ALN_DTYPE, count=tot)
    E.emu_free(pool)
    return n_aln, max_entries, aln_off, aln, stats


_wemu = {}


def wemu(team: int = 1):
    """tests/host_emu/warp_emu.cpp: the body of k_search_warp on `team` warps of 32 coroutine lanes (team > 1: the warps of a block
    share one read)."""
    if team not in _wemu:
        so = os.path.join(EMU_DIR, "libwarp_emu.so" if team == 1 else f"libwarp_emu_team{team}.so")
        srcs = [os.path.join(EMU_DIR, "warp_emu.cpp"), os.path.join(EMU_DIR, "host_emu_shim.h")] + [
            os.path.join(ROOT, "network-aware-bwa_b200", "csrc", f) for f in ("search_warp.cuh", "kernels.cuh", "fmindex.cuh", "hostprep.h")]
        if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
            subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-w", f"-DEMU_TEAM={team}", "-I", EMU_DIR, "-o", so, srcs[0]], check=True)
        E = C.CDLL(so)
        E.wemu_last_error.restype = C.c_char_p
        E.wemu_load_index.argtypes = [C.POINTER(C.POINTER(abi.bwt_t))]
        E.wemu_load_index.restype = C.c_void_p
        E.wemu_free_index.argtypes = [C.c_void_p]
        E.wemu_aln_flat.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(abi.gap_opt_t), C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.POINTER(C.c_void_p), C.c_uint32, C.POINTER(C.c_int)]
        E.wemu_free.argtypes = [C.c_void_p]
        _wemu[team] = E
    return _wemu[team]


def wemu_aln(h, reads, opt, pool_chunks=4096, team=1):
    """-> (n_aln, max_entries, aln_off, aln, reads that found the chunk pool dry)"""
    E = wemu(team)
    n = reads.n
    bases = np.ascontiguousarray(reads.bases, dtype=np.uint8)
    offs = np.ascontiguousarray(reads.offs, dtype=np.int64)
    n_aln = np.empty(n, dtype=np.int32)
    max_entries = np.zeros(n, dtype=np.int32)
    aln_off = np.empty(n + 1, dtype=np.int64)
    pool = C.c_void_p()
    dry = C.c_int(0)
    rc = E.wemu_aln_flat(h, n, bases.ctypes.data, offs.ctypes.data, C.byref(opt), n_aln.ctypes.data, max_entries.ctypes.data,
                         aln_off.ctypes.data, C.byref(pool), pool_chunks, C.byref(dry))
    if rc:
        raise RuntimeError(E.wemu_last_error().decode())
    tot = int(aln_off[n])
    aln = np.empty(tot, dtype=abi.ALN_DTYPE)
    if tot:
        buf = (C.c_char * (16 * tot)).from_address(pool.value)
        aln[:] = np.frombuffer(buf, dtype=abi.ALN_DTYPE, count=tot)
    E.wemu_free(pool)
    return n_aln, max_entries, aln_off, aln, dry.value


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


# ---------------------------------------------------------------- this repo's C restatement (oracle/liboracle.so)
class orc_index_t(C.Structure):
    _fields_ = [("primary", C.c_uint32), ("seq_len", C.c_uint32), ("L2", C.c_uint32 * 5),
                ("bwt", C.POINTER(C.c_uint32)), ("sa", C.POINTER(C.c_uint32)), ("sa_intv", C.c_int)]


class orc_opt_t(C.Structure):
    _fields_ = [("s_mm", C.c_int), ("s_gapo", C.c_int), ("s_gape", C.c_int), ("mode", C.c_int),
                ("indel_end_skip", C.c_int), ("max_del_occ", C.c_int), ("max_entries", C.c_int),
                ("fnr", C.c_double),
                ("max_diff", C.c_int), ("max_gapo", C.c_int), ("max_gape", C.c_int), ("max_seed_diff", C.c_int),
                ("seed_len", C.c_int), ("max_top2", C.c_int)]


_orc = None


def orc():
    global _orc
    if _orc is None:
        so = os.path.join(ROOT, "oracle", "liboracle.so")
        src = [os.path.join(ROOT, "oracle", f) for f in ("bwa_oracle.c", "bwa_oracle.h")]
        if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
            subprocess.run(["gcc", "-O2", "-Wall", "-fPIC", "-shared", "-o", so, src[0], "-lm"], check=True)
        O = C.CDLL(so)
        O.orc_occ.argtypes = [C.POINTER(orc_index_t), C.c_uint32, C.c_int]
        O.orc_occ.restype = C.c_uint32
        O.orc_sa.argtypes = [C.POINTER(orc_index_t), C.c_uint32]
        O.orc_sa.restype = C.c_uint32
        O.orc_cal_maxdiff.argtypes = [C.c_int, C.c_double, C.c_double]
        O.orc_aln_flat.argtypes = [C.POINTER(orc_index_t), C.c_int, C.c_void_p, C.c_void_p, C.POINTER(orc_opt_t),
                                   C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]
        O.orc_free.argtypes = [C.c_void_p]
        O.orc_sw_local.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        _orc = O
    return _orc


def orc_index(idx):
    arr = (orc_index_t * 2)()
    for s in range(2):
        b = idx.bwt[s]
        arr[s].primary, arr[s].seq_len, arr[s].sa_intv = b.primary, b.seq_len, b.sa_intv
        for i in range(5):
            arr[s].L2[i] = int(b.L2[i])
        arr[s].bwt = b.bwt.ctypes.data_as(C.POINTER(C.c_uint32))
        arr[s].sa = b.sa.ctypes.data_as(C.POINTER(C.c_uint32))
    return arr


def orc_opt(opt) -> orc_opt_t:
    o = orc_opt_t()
    for f, _ in orc_opt_t._fields_:
        setattr(o, f, getattr(opt, f))
    return o


def orc_aln(oidx, reads, opt):
    O = orc()
    n = reads.n
    bases = np.ascontiguousarray(reads.bases, dtype=np.uint8)
    offs = np.ascontiguousarray(reads.offs, dtype=np.int64)
    n_aln = np.empty(n, dtype=np.int32)
    max_entries = np.zeros(n, dtype=np.int32)
    pool = C.c_void_p()
    oo = orc_opt(opt)
    O.orc_aln_flat(oidx, n, bases.ctypes.data, offs.ctypes.data, C.byref(oo), n_aln.ctypes.data, max_entries.ctypes.data,
                   C.byref(pool))
    aln_off = np.zeros(n + 1, dtype=np.int64)
    aln_off[1:] = np.cumsum(n_aln)
    tot = int(aln_off[n])
    aln = np.empty(tot, dtype=abi.ALN_DTYPE)
    if tot:
        buf = (C.c_char * (16 * tot)).from_address(pool.value)
        aln[:] = np.frombuffer(buf, dtype=abi.ALN_DTYPE, count=tot)
    O.orc_free(pool)
    return n_aln, max_entries, aln_off, aln


def orc_sw_batch(refs, ref_off, queries, q_off) -> np.ndarray:
    """-> int32[n,5]: score, start_i, start_j, end_i, end_j"""
    O = orc()
    n = ref_off.size - 1
    out = np.zeros((n, 5), dtype=np.int32)
    res = (C.c_int * 4)()
    for i in range(n):
        r = np.ascontiguousarray(refs[ref_off[i]:ref_off[i + 1]])
        q = np.ascontiguousarray(queries[q_off[i]:q_off[i + 1]])
        out[i, 0] = O.orc_sw_local(r.ctypes.data, r.size, q.ctypes.data, q.size, res)
        out[i, 1:] = list(res)
    return out


def ref_sw_batch(refs, ref_off, queries, q_off, threads: int = 8) -> np.ndarray:
    _, H = ref()
    H.refh_sw_batch.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    refs = np.ascontiguousarray(refs, dtype=np.uint8)
    queries = np.ascontiguousarray(queries, dtype=np.uint8)
    ref_off = np.ascontiguousarray(ref_off, dtype=np.int64)
    q_off = np.ascontiguousarray(q_off, dtype=np.int64)
    n = ref_off.size - 1
    out = np.zeros((n, 5), dtype=np.int32)
    H.refh_sw_batch(n, refs.ctypes.data, ref_off.ctypes.data, queries.ctypes.data, q_off.ctypes.data, out.ctypes.data, threads)
    return out


def make_sw_jobs(T, n_jobs: int, seed: int, read_len=(30, 120), win=(60, 420), ref_n=True, with_beg=False):
    """Mate-rescue-like SW jobs: a reference window and a read that (mostly) comes from it,
    with substitutions, an indel, N, or nothing in common.  -> refs, ref_off, queries, q_off"""
    rng = np.random.default_rng(seed)
    refs, queries, ro, qo, begs = [], [], [0], [0], []
    n = T.size
    for j in range(n_jobs):
        L = int(rng.integers(read_len[0], read_len[1] + 1))
        W = int(rng.integers(max(win[0], 20), win[1] + 1))
        beg = int(rng.integers(0, n - W - 1))
        window = T[beg:beg + W].copy()
        kind = j % 8
        if kind == 7:  # unrelated read
            q = rng.integers(0, 4, size=L, dtype=np.uint8)
        else:
            L = min(L, W)
            s = int(rng.integers(0, W - L + 1))
            q = window[s:s + L].copy()
            sub = rng.random(L) < (0.02 if kind < 4 else 0.10)
            q[sub] = (q[sub] + rng.integers(1, 4, size=int(sub.sum()), dtype=np.uint8)) & 3
            if kind in (2, 5) and L > 20:  # deletion from the read
                p, d = int(rng.integers(5, L - 10)), int(rng.integers(1, 6))
                q = np.concatenate([q[:p], q[p + d:]])
            if kind in (3, 6) and L > 20:  # insertion into the read
                p, d = int(rng.integers(5, L - 10)), int(rng.integers(1, 6))
                q = np.concatenate([q[:p], rng.integers(0, 4, size=d, dtype=np.uint8), q[p:]])
            if kind == 4:
                q[rng.random(q.size) < 0.03] = 4
        if ref_n and j % 11 == 0:
            window[rng.random(W) < 0.02] = 4
        begs.append(beg)
        refs.append(window); queries.append(q.astype(np.uint8))
        ro.append(ro[-1] + window.size); qo.append(qo[-1] + q.size)
    out = (np.concatenate(refs).astype(np.uint8), np.array(ro, dtype=np.int64),
           np.concatenate(queries).astype(np.uint8), np.array(qo, dtype=np.int64))
    return out + (np.array(begs, dtype=np.int64),) if with_beg else out


def path_to_cigar(path_ijc: np.ndarray) -> np.ndarray:
    """aln_path2cigar32 + bwa_aln_path2cigar (stdaln.c:1009-1039, bwtaln.c:396-406) on a (n,3) path."""
    if path_ijc.shape[0] == 0:
        return np.empty(0, dtype=np.uint16)
    t = path_ijc[::-1, 2]
    cut = np.nonzero(np.diff(t) != 0)[0] + 1
    starts = np.concatenate([[0], cut]); ends = np.concatenate([cut, [t.size]])
    return ((t[starts].astype(np.uint16) << 14) | (ends - starts).astype(np.uint16)).astype(np.uint16)


def ref_sw_path(ref_arr: np.ndarray, q_arr: np.ndarray):
    """aln_local_core with path (as bwa_sw_core calls it) -> (score, start_i, start_j, end_i, end_j, cigar)"""
    _, H = ref()
    r = np.ascontiguousarray(ref_arr, dtype=np.uint8); q = np.ascontiguousarray(q_arr, dtype=np.uint8)
    path = np.zeros(3 * (r.size + q.size + 4), dtype=np.int32)
    plen = C.c_int()
    score = H.refh_sw1(r.ctypes.data, r.size, q.ctypes.data, q.size, C.byref(plen), path.ctypes.data)
    p = path[:3 * plen.value].reshape(-1, 3)
    if plen.value == 0:
        return (score, 0, 0, 0, 0, np.empty(0, dtype=np.uint16))
    return (score, int(p[-1, 0]), int(p[-1, 1]), int(p[0, 0]), int(p[0, 1]), path_to_cigar(p))


def ref_global(ref_arr: np.ndarray, q_arr: np.ndarray, gap_end: int, band: int):
    _, H = ref()
    H.refh_global.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_void_p]
    r = np.ascontiguousarray(ref_arr, dtype=np.uint8); q = np.ascontiguousarray(q_arr, dtype=np.uint8)
    path = np.zeros(3 * (r.size + q.size + 4), dtype=np.int32)
    plen = C.c_int()
    score = H.refh_global(r.ctypes.data, r.size, q.ctypes.data, q.size, gap_end, band, C.byref(plen), path.ctypes.data)
    p = path[:3 * plen.value].reshape(-1, 3)
    if plen.value == 0:
        return (score, 0, 0, 0, 0, np.empty(0, dtype=np.uint16))
    return (score, int(p[-1, 0]), int(p[-1, 1]), int(p[0, 0]), int(p[0, 1]), path_to_cigar(p))
