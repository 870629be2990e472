"""Python binding of the C-ABI (include/bwa_gpu.h) -- ctypes over libbwagpu.so.

This is the stub a host in another language would write (INTEGRATION.md shows the C
version for bam2bam.c).  It adds nothing: every call goes straight to the library, and
there is no fallback of any kind -- if the library is missing, cannot be loaded or finds no
sm_100 device, the call raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import abi
from .build import LIB


class BwaGpuError(RuntimeError):
    pass


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        path = os.environ.get("BWAGPU_LIB", LIB)  # experiment builds (build.build_variant)
        if not os.path.exists(path):
            raise BwaGpuError(f"{path} is not built (run __graft_entry__.build()); there is no CPU fallback")
        L = C.CDLL(path)
        L.bwa_gpu_last_error.restype = C.c_char_p
        L.bwa_gpu_init.argtypes = [C.c_int, C.POINTER(C.c_int)]
        L.bwa_gpu_load_index.argtypes = [C.POINTER(C.POINTER(abi.bwt_t)), C.c_void_p, C.c_int64]
        L.bwa_gpu_destroy.restype = None
        L.bwa_gpu_cal_sa_reads_gap.argtypes = [C.c_int, C.POINTER(abi.bwa_seq_t), C.POINTER(abi.gap_opt_t)]
        L.bwa_gpu_free_alns.argtypes = [C.c_int, C.POINTER(abi.bwa_seq_t)]
        L.bwa_gpu_free_alns.restype = None
        L.bwa_gpu_aln_flat.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.POINTER(abi.gap_opt_t), C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]
        L.bwa_gpu_cal_pac_pos.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.bwa_gpu_mate_sw.argtypes = [C.c_int, C.POINTER(abi.sw_job_t), C.POINTER(abi.sw_res_t)]
        L.bwa_gpu_mate_sw_path.argtypes = [C.c_int, C.POINTER(abi.sw_job_t), C.POINTER(abi.path_res_t), C.POINTER(C.c_void_p)]
        L.bwa_gpu_global_align.argtypes = [C.c_int, C.POINTER(abi.sw_job_t), C.c_int, C.c_int, C.POINTER(abi.path_res_t),
                                           C.POINTER(C.c_void_p)]
        L.bwa_gpu_global_align_seqs.argtypes = [C.c_int, C.POINTER(abi.ga_job_t), C.c_int, C.c_int, C.POINTER(abi.path_res_t),
                                                C.POINTER(C.c_void_p)]
        L.bwa_gpu_get_stats.argtypes = [C.POINTER(abi.stats_t)]
        L.bwa_gpu_get_totals.argtypes = [C.POINTER(abi.totals_t)]
        L.bwa_gpu_reset_totals.restype = None
        L.bwa_gpu_set_stats.argtypes = [C.c_int]
        L.bwa_gpu_resident_stage.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.POINTER(abi.gap_opt_t)]
        L.bwa_gpu_resident_run.argtypes = [C.POINTER(C.c_double)]
        L.bwa_gpu_resident_fetch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]
        L.bwa_gpu_index_build.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.POINTER(abi.bwt_t), C.POINTER(abi.bwt_t)]
        L.bwa_gpu_index_free.argtypes = [C.POINTER(abi.bwt_t)]
        L.bwa_gpu_index_free.restype = None
        L.bwa_gpu_index_write.argtypes = [C.c_char_p, C.POINTER(abi.bwt_t), C.POINTER(abi.bwt_t)]
        L.bwa_gpu_bgzf_deflate.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_int64),
                                           C.POINTER(C.c_void_p), C.POINTER(C.c_int32), C.POINTER(C.c_double)]
        L.bwa_gpu_bgzf_inflate.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.POINTER(C.c_double)]
        L.bwa_gpu_host_alloc.argtypes = [C.c_size_t]
        L.bwa_gpu_host_alloc.restype = C.c_void_p
        L.bwa_gpu_host_free.argtypes = [C.c_void_p]
        L.bwa_gpu_host_free.restype = None
        _lib = L
    return _lib


EXPORTS = [
    "bwa_gpu_init", "bwa_gpu_load_index", "bwa_gpu_load_pac", "bwa_gpu_destroy", "bwa_gpu_last_error",
    "bwa_gpu_cal_sa_reads_gap", "bwa_gpu_free_alns", "bwa_gpu_aln_flat", "bwa_gpu_cal_pac_pos", "bwa_gpu_mate_sw", "bwa_gpu_mate_sw_path", "bwa_gpu_global_align",
    "bwa_gpu_global_align_seqs",
    "bwa_gpu_get_stats", "bwa_gpu_set_stats", "bwa_gpu_get_totals", "bwa_gpu_reset_totals", "bwa_gpu_probe_random_sectors",
    "bwa_gpu_resident_stage", "bwa_gpu_resident_run", "bwa_gpu_resident_fetch",
    "bwa_gpu_index_build", "bwa_gpu_index_free", "bwa_gpu_index_write",
    "bwa_gpu_bgzf_deflate", "bwa_gpu_bgzf_inflate", "bwa_gpu_host_alloc", "bwa_gpu_host_free",
]


def _ck(rc: int) -> None:
    if rc != 0:
        raise BwaGpuError(lib().bwa_gpu_last_error().decode())


def init(device_ids=None) -> None:
    if device_ids is None:
        _ck(lib().bwa_gpu_init(0, None))
    else:
        arr = (C.c_int * len(device_ids))(*device_ids)
        _ck(lib().bwa_gpu_init(len(device_ids), arr))


def destroy() -> None:
    lib().bwa_gpu_destroy()


def load_index(idx) -> None:
    """idx: index.FMIndex (host arrays in the reference's layout)."""
    t0, t1 = abi.make_bwt_t(idx.bwt[0]), abi.make_bwt_t(idx.bwt[1])
    arr = (C.POINTER(abi.bwt_t) * 2)(C.pointer(t0), C.pointer(t1))
    pac = np.ascontiguousarray(idx.pac)
    _ck(lib().bwa_gpu_load_index(arr, pac.ctypes.data, idx.l_pac))


def index_build(pac: np.ndarray, l_pac: int, device: int = 0, write_prefix: str | None = None):
    """bwa_gpu_index_build: packed forward sequence -> (fwd, rev) as index.Bwt objects (numpy copies of the
    reference-layout arrays).  write_prefix: also dump <prefix>.bwt/.rbwt/.sa/.rsa through bwa_gpu_index_write."""
    from . import index as ix
    pac = np.ascontiguousarray(pac, dtype=np.uint8)
    if pac.size < l_pac // 4 + 1:
        raise ValueError("pac shorter than l_pac/4+1 bytes")
    t = [abi.bwt_t(), abi.bwt_t()]
    _ck(lib().bwa_gpu_index_build(pac.ctypes.data, int(l_pac), int(device), C.byref(t[0]), C.byref(t[1])))
    try:
        if write_prefix is not None:
            _ck(lib().bwa_gpu_index_write(write_prefix.encode(), C.byref(t[0]), C.byref(t[1])))
        out = []
        for b in t:
            L2 = np.array([b.L2[i] for i in range(5)], dtype=np.uint32)
            bwt = np.ctypeslib.as_array(b.bwt, shape=(b.bwt_size,)).copy()
            sa = np.ctypeslib.as_array(b.sa, shape=(b.n_sa,)).copy()
            out.append(ix.Bwt(primary=int(b.primary), L2=L2, seq_len=int(b.seq_len), bwt=bwt, sa=sa, sa_intv=int(b.sa_intv)))
    finally:
        for b in t:
            lib().bwa_gpu_index_free(C.byref(b))
    return out


def cal_sa_reads_gap(seqs, opt) -> None:
    """seqs: ctypes array of abi.bwa_seq_t; filled in place like bwa_cal_sa_reg_gap."""
    _ck(lib().bwa_gpu_cal_sa_reads_gap(len(seqs), seqs, C.byref(opt)))


def _wrap_pool(ptr, n):
    if n == 0:
        return np.empty(0, dtype=abi.ALN_DTYPE)
    buf = (C.c_char * (16 * n)).from_address(ptr.value)
    return np.frombuffer(buf, dtype=abi.ALN_DTYPE, count=n).copy()


def aln_flat(bases: np.ndarray, offs: np.ndarray, opt):
    """-> (n_aln int32[n], max_entries int32[n], aln_off int64[n+1], aln structured[n_total])"""
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    offs = np.ascontiguousarray(offs, dtype=np.int64)
    n = offs.size - 1
    n_aln = np.empty(n, dtype=np.int32)
    max_entries = np.empty(n, dtype=np.int32)
    aln_off = np.empty(n + 1, dtype=np.int64)
    pool = C.c_void_p()
    _ck(lib().bwa_gpu_aln_flat(n, bases.ctypes.data, offs.ctypes.data, C.byref(opt), n_aln.ctypes.data,
                               max_entries.ctypes.data, aln_off.ctypes.data, C.byref(pool)))
    return n_aln, max_entries, aln_off, _wrap_pool(pool, int(aln_off[n]))


def cal_pac_pos(sa_idx: np.ndarray, which: np.ndarray) -> np.ndarray:
    sa_idx = np.ascontiguousarray(sa_idx, dtype=np.uint32)
    which = np.ascontiguousarray(which, dtype=np.uint8)
    out = np.empty(sa_idx.size, dtype=np.uint32)
    _ck(lib().bwa_gpu_cal_pac_pos(sa_idx.size, sa_idx.ctypes.data, which.ctypes.data, out.ctypes.data))
    return out


def mate_sw(jobs):
    """jobs: list of (beg, reglen, seq uint8 array) -> list of (score, start_i, start_j, end_i, end_j)"""
    n = len(jobs)
    arr = (abi.sw_job_t * n)()
    keep = []
    for i, (beg, reglen, seq) in enumerate(jobs):
        s = np.ascontiguousarray(seq, dtype=np.uint8)
        keep.append(s)
        arr[i].beg, arr[i].reglen, arr[i].len = beg, reglen, s.size
        arr[i].seq = s.ctypes.data_as(C.POINTER(C.c_ubyte))
    res = (abi.sw_res_t * n)()
    _ck(lib().bwa_gpu_mate_sw(n, arr, res))
    return [(r.score, r.start_i, r.start_j, r.end_i, r.end_j) for r in res]


def _sw_jobs(jobs):
    n = len(jobs)
    arr = (abi.sw_job_t * n)()
    keep = []
    for i, (beg, reglen, seq) in enumerate(jobs):
        s = np.ascontiguousarray(seq, dtype=np.uint8)
        keep.append(s)
        arr[i].beg, arr[i].reglen, arr[i].len = beg, reglen, s.size
        arr[i].seq = s.ctypes.data_as(C.POINTER(C.c_ubyte))
    return arr, keep


def _path_results(res, pool, n):
    out = []
    for i in range(n):
        r = res[i]
        cig = np.empty(0, dtype=np.uint16)
        if r.n_cigar:
            buf = (C.c_uint16 * r.n_cigar).from_address(pool.value + 2 * r.cigar_off)
            cig = np.frombuffer(buf, dtype=np.uint16).copy()
        out.append((r.score, r.start_i, r.start_j, r.end_i, r.end_j, cig))
    return out


def mate_sw_path(jobs):
    """jobs as for mate_sw -> list of (score, start_i, start_j, end_i, end_j, cigar uint16[])"""
    n = len(jobs)
    arr, keep = _sw_jobs(jobs)
    res = (abi.path_res_t * n)()
    pool = C.c_void_p()
    _ck(lib().bwa_gpu_mate_sw_path(n, arr, res, C.byref(pool)))
    return _path_results(res, pool, n)


def global_align(jobs, gap_end: int = 5, band: int = 50):
    n = len(jobs)
    arr, keep = _sw_jobs(jobs)
    res = (abi.path_res_t * n)()
    pool = C.c_void_p()
    _ck(lib().bwa_gpu_global_align(n, arr, gap_end, band, res, C.byref(pool)))
    return _path_results(res, pool, n)


def global_align_seqs(pairs, gap_end: int = 5, band: int = 50):
    """pairs: list of (ref uint8 array, seq uint8 array) -> like global_align"""
    n = len(pairs)
    arr = (abi.ga_job_t * n)()
    keep = []
    for i, (ref, seq) in enumerate(pairs):
        r = np.ascontiguousarray(ref, dtype=np.uint8)
        s = np.ascontiguousarray(seq, dtype=np.uint8)
        keep += [r, s]
        arr[i].ref = r.ctypes.data_as(C.POINTER(C.c_ubyte))
        arr[i].reflen = r.size
        arr[i].seq = s.ctypes.data_as(C.POINTER(C.c_ubyte))
        arr[i].len = s.size
    res = (abi.path_res_t * n)()
    pool = C.c_void_p()
    _ck(lib().bwa_gpu_global_align_seqs(n, arr, gap_end, band, res, C.byref(pool)))
    return _path_results(res, pool, n)


def bgzf_deflate(data, level: int = 2):
    """bwa_gpu_bgzf_deflate: bytes of BAM stream -> (the BGZF members back to back, member sizes, kernel ms)."""
    buf = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else np.ascontiguousarray(data, dtype=np.uint8)
    out, out_n, lens, n_mem, ms = C.c_void_p(), C.c_int64(), C.c_void_p(), C.c_int32(), C.c_double()
    _ck(lib().bwa_gpu_bgzf_deflate(buf.ctypes.data if buf.size else None, buf.size, level, C.byref(out), C.byref(out_n), C.byref(lens),
                                   C.byref(n_mem), C.byref(ms)))
    packed = C.string_at(out.value, out_n.value) if out_n.value else b""
    member_len = np.ctypeslib.as_array(C.cast(lens.value, C.POINTER(C.c_int32)), shape=(n_mem.value,)).copy() if n_mem.value else np.zeros(0, np.int32)
    return packed, member_len, ms.value


def bgzf_member_offsets(data: bytes) -> np.ndarray:
    """Start of every BGZF member of `data` (+ its end): the BSIZE chain of the BC extra fields (bgzf.c:274-291)."""
    offs, p = [0], 0
    while p < len(data):
        assert data[p:p + 4] == b"\x1f\x8b\x08\x04" and data[p + 12:p + 14] == b"BC", f"not a BGZF member at {p}"
        p += int.from_bytes(data[p + 16:p + 18], "little") + 1
        offs.append(p)
    return np.array(offs, dtype=np.int64)


def bgzf_inflate(data: bytes):
    """bwa_gpu_bgzf_inflate on a whole BGZF byte string -> (decompressed bytes, per-member output offsets, kernel ms)."""
    buf = np.frombuffer(data, dtype=np.uint8)
    moff = bgzf_member_offsets(data)
    n = moff.size - 1
    cap = 65536 * max(1, n)
    out = np.empty(cap, dtype=np.uint8)
    ooff = np.zeros(n + 1, dtype=np.int64)
    ms = C.c_double()
    _ck(lib().bwa_gpu_bgzf_inflate(buf.ctypes.data, buf.size, n, moff.ctypes.data, out.ctypes.data, cap, ooff.ctypes.data, C.byref(ms)))
    return out[: ooff[n]].tobytes(), ooff, ms.value


def set_stats(enabled: bool) -> None:
    _ck(lib().bwa_gpu_set_stats(1 if enabled else 0))


def probe_random_sectors(buffer_bytes: int, chains: int = 4, steps: int = 256) -> float:
    """GB/s of dependent random 32-byte sector loads over a buffer of that size (roofline denominator)."""
    out = C.c_double(0.0)
    L = lib()
    L.bwa_gpu_probe_random_sectors.argtypes = [C.c_int64, C.c_int, C.c_int, C.POINTER(C.c_double)]
    _ck(L.bwa_gpu_probe_random_sectors(int(buffer_bytes), int(chains), int(steps), C.byref(out)))
    return out.value


def get_stats() -> dict:
    s = abi.stats_t()
    _ck(lib().bwa_gpu_get_stats(C.byref(s)))
    return s.asdict()


def get_totals() -> dict:
    t = abi.totals_t()
    _ck(lib().bwa_gpu_get_totals(C.byref(t)))
    return t.asdict()


def reset_totals() -> None:
    lib().bwa_gpu_reset_totals()


def resident_stage(bases: np.ndarray, offs: np.ndarray, opt) -> None:
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    offs = np.ascontiguousarray(offs, dtype=np.int64)
    _ck(lib().bwa_gpu_resident_stage(offs.size - 1, bases.ctypes.data, offs.ctypes.data, C.byref(opt)))


def resident_run() -> float:
    ms = C.c_double()
    _ck(lib().bwa_gpu_resident_run(C.byref(ms)))
    return ms.value


def resident_fetch(n: int):
    n_aln = np.empty(n, dtype=np.int32)
    max_entries = np.empty(n, dtype=np.int32)
    aln_off = np.empty(n + 1, dtype=np.int64)
    pool = C.c_void_p()
    _ck(lib().bwa_gpu_resident_fetch(n_aln.ctypes.data, max_entries.ctypes.data, aln_off.ctypes.data, C.byref(pool)))
    return n_aln, max_entries, aln_off, _wrap_pool(pool, int(aln_off[n]))
