"""ctypes mirrors of the reference structs that cross the C-ABI (include/bwa_gpu.h).

Layouts follow bwt.h:43-59 (bwt_t), bwtaln.h:43-47 (bwt_aln1_t), 58-62 (bwt_multi1_t),
64-90 (bwa_seq_t), 143-153 (gap_opt_t); sizes are asserted against the reference's own
headers in tests/test_abi.py (golden offsets in tests/golden/abi_layout.json).
The same classes are used to call the reference itself (oracle/_ref/libbwaref.so) in the
parity tests, so both sides see identical bytes.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

bwtint_t = C.c_uint32


class bwt_t(C.Structure):
    _fields_ = [
        ("primary", bwtint_t),
        ("L2", bwtint_t * 5),
        ("seq_len", bwtint_t),
        ("bwt_size", bwtint_t),
        ("bwt", C.POINTER(C.c_uint32)),
        ("cnt_table", C.c_uint32 * 256),
        ("sa_intv", C.c_int),
        ("n_sa", bwtint_t),
        ("sa", C.POINTER(bwtint_t)),
    ]


class bwt_aln1_t(C.Structure):
    _fields_ = [
        ("n_mm", C.c_uint32, 8), ("n_gapo", C.c_uint32, 8), ("n_gape", C.c_uint32, 8), ("a", C.c_uint32, 1),
        ("k", bwtint_t), ("l", bwtint_t), ("score", C.c_int),
    ]


ALN_DTYPE = np.dtype([("info", "<u4"), ("k", "<u4"), ("l", "<u4"), ("score", "<i4")])
"""numpy view of bwt_aln1_t: info = n_mm | n_gapo << 8 | n_gape << 16 | a << 24."""


class bwt_multi1_t(C.Structure):
    _fields_ = [
        ("pos", C.c_uint32),
        ("n_cigar", C.c_uint32, 15), ("gap", C.c_uint32, 8), ("mm", C.c_uint32, 8), ("strand", C.c_uint32, 1),
        ("cigar", C.POINTER(C.c_uint16)),
    ]


class bwa_seq_t(C.Structure):
    _fields_ = [
        ("name", C.c_void_p),
        ("seq", C.POINTER(C.c_ubyte)), ("rseq", C.POINTER(C.c_ubyte)), ("qual", C.POINTER(C.c_ubyte)),
        ("len", C.c_uint32, 20), ("strand", C.c_uint32, 1), ("type", C.c_uint32, 2), ("dummy", C.c_uint32, 1),
        ("extra_flag", C.c_uint32, 8),
        ("n_mm", C.c_uint32, 8), ("n_gapo", C.c_uint32, 8), ("n_gape", C.c_uint32, 8), ("mapQ", C.c_uint32, 8),
        ("score", C.c_int),
        ("clip_len", C.c_int),
        ("n_aln", C.c_int),
        ("aln", C.POINTER(bwt_aln1_t)),
        ("n_multi", C.c_int),
        ("multi", C.POINTER(bwt_multi1_t)),
        ("sa", bwtint_t), ("pos", bwtint_t),
        ("c1", C.c_uint64, 28), ("c2", C.c_uint64, 28), ("seQ", C.c_uint64, 8),
        ("n_cigar", C.c_int),
        ("cigar", C.POINTER(C.c_uint16)),
        ("tid", C.c_int),
        ("bc", C.c_char * 64),
        ("full_len", C.c_uint32, 20), ("nm", C.c_uint32, 12),
        ("md", C.c_void_p),
        ("max_entries", C.c_int),
    ]


class gap_opt_t(C.Structure):
    _fields_ = [
        ("s_mm", C.c_int), ("s_gapo", C.c_int), ("s_gape", C.c_int),
        ("mode", C.c_int),
        ("indel_end_skip", C.c_int), ("max_del_occ", C.c_int), ("max_entries", C.c_int),
        ("fnr", C.c_float),
        ("max_diff", C.c_int), ("max_gapo", C.c_int), ("max_gape", C.c_int),
        ("max_seed_diff", C.c_int), ("seed_len", C.c_int),
        ("n_threads", C.c_int),
        ("max_top2", C.c_int),
        ("trim_qual", C.c_int),
    ]


BWA_MODE_GAPE = 0x01
BWA_MODE_COMPREAD = 0x02
BWA_MODE_LOGGAP = 0x04
BWA_MODE_NONSTOP = 0x10


def default_gap_opt(**kw) -> gap_opt_t:
    """gap_init_opt (bwtaln.c:19-35)."""
    o = gap_opt_t()
    o.s_mm, o.s_gapo, o.s_gape = 3, 11, 4
    o.max_diff, o.max_gapo, o.max_gape = -1, 1, 6
    o.indel_end_skip, o.max_del_occ, o.max_entries = 5, 10, 2000000
    o.mode = BWA_MODE_GAPE | BWA_MODE_COMPREAD
    o.seed_len, o.max_seed_diff = 32, 2
    o.fnr = 0.04
    o.n_threads = 1
    o.max_top2 = 30
    o.trim_qual = 0
    for k, v in kw.items():
        if not hasattr(o, k):
            raise AttributeError(k)
        setattr(o, k, v)
    return o


class sw_job_t(C.Structure):
    _fields_ = [("beg", C.c_int64), ("reglen", C.c_int32), ("len", C.c_int32), ("seq", C.POINTER(C.c_ubyte))]


class ga_job_t(C.Structure):
    _fields_ = [("ref", C.POINTER(C.c_ubyte)), ("reflen", C.c_int32), ("len", C.c_int32), ("seq", C.POINTER(C.c_ubyte))]


class sw_res_t(C.Structure):
    _fields_ = [("score", C.c_int32), ("start_i", C.c_int32), ("start_j", C.c_int32),
                ("end_i", C.c_int32), ("end_j", C.c_int32)]


class path_res_t(C.Structure):
    _fields_ = [("score", C.c_int32), ("n_cigar", C.c_int32), ("start_i", C.c_int32), ("start_j", C.c_int32),
                ("end_i", C.c_int32), ("end_j", C.c_int32), ("cigar_off", C.c_int64)]


class stats_t(C.Structure):
    _fields_ = [
        ("ms_h2d", C.c_double), ("ms_width", C.c_double), ("ms_search", C.c_double), ("ms_compact", C.c_double),
        ("ms_d2h", C.c_double), ("ms_total_device", C.c_double), ("ms_host_marshal", C.c_double),
        ("n_reads", C.c_int64), ("n_aln", C.c_int64), ("n_overflow_t2", C.c_int64), ("n_overflow_t3", C.c_int64),
        ("occ_fetches_width", C.c_int64), ("occ_fetches_search", C.c_int64),
        ("own_fetches_width", C.c_int64), ("own_fetches_search", C.c_int64),
        ("n_pops", C.c_int64), ("n_pushes", C.c_int64),
        ("launches", C.c_int32), ("n_devices", C.c_int32),
        ("ms_tier", C.c_double * 4),
        ("n_stored", C.c_int64),
        ("n_pruned", C.c_int64), ("n_expand", C.c_int64), ("n_exact", C.c_int64), ("n_derive", C.c_int64),
        ("n_trips", C.c_int64), ("ms_sw_kernel", C.c_double), ("x_chunks_used", C.c_int64), ("ns_queue_empty", C.c_int64), ("ns_kernel", C.c_int64),
    ]

    def asdict(self) -> dict:
        d = {f[0]: getattr(self, f[0]) for f in self._fields_}
        d["ms_tier"] = list(self.ms_tier)
        return d


class totals_t(C.Structure):
    _fields_ = [
        ("ms_width", C.c_double), ("ms_search", C.c_double), ("ms_sa", C.c_double), ("ms_sw", C.c_double), ("ms_global", C.c_double),
        ("ms_search_pass", C.c_double * 3),
        ("launches", C.c_int64),
        ("reads", C.c_int64), ("alns", C.c_int64), ("sa_queries", C.c_int64), ("sw_jobs", C.c_int64), ("ga_jobs", C.c_int64),
        ("sw_cells_fwd", C.c_int64),
        ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64),
        ("occ_fetches_width", C.c_int64), ("occ_fetches_search", C.c_int64), ("own_fetches_search", C.c_int64),
        ("ms_bgzf", C.c_double), ("bgzf_bytes_in", C.c_int64), ("bgzf_bytes_out", C.c_int64),
        ("ms_inflate", C.c_double), ("inflate_bytes_in", C.c_int64), ("inflate_bytes_out", C.c_int64),
    ]

    def asdict(self) -> dict:
        d = {f[0]: getattr(self, f[0]) for f in self._fields_}
        d["ms_search_pass"] = list(self.ms_search_pass)
        return d


def make_bwt_t(b) -> bwt_t:
    """bwt_t view over an index.Bwt (arrays stay owned by the Bwt object; keep it alive)."""
    t = bwt_t()
    t.primary = b.primary
    for i in range(5):
        t.L2[i] = int(b.L2[i])
    t.seq_len = b.seq_len
    t.bwt_size = b.bwt.size
    t.bwt = b.bwt.ctypes.data_as(C.POINTER(C.c_uint32))
    for i in range(256):  # bwt_gen_cnt_table (bwt.c:36-45); only the CPU reference reads it
        x = 0
        for j in range(4):
            x |= (((i & 3) == j) + ((i >> 2 & 3) == j) + ((i >> 4 & 3) == j) + ((i >> 6) == j)) << (j << 3)
        t.cnt_table[i] = x
    t.sa_intv = b.sa_intv
    t.n_sa = b.n_sa
    t.sa = b.sa.ctypes.data_as(C.POINTER(bwtint_t)) if b.sa is not None else None
    return t


def make_seqs(reads, idxs=None):
    """bwa_seq_t array for reads (simulate.Reads) the way bam1_to_seq fills it
    (bwaseqio.c:272-297): seq = reversed read, rseq = reverse complement.  Returns
    (array, keepalive)."""
    idxs = range(reads.n) if idxs is None else idxs
    n = len(idxs)
    arr = (bwa_seq_t * n)()
    keep = []
    for j, i in enumerate(idxs):
        r = np.ascontiguousarray(reads.read(i))
        s0 = np.ascontiguousarray(r[::-1])
        s1 = np.where(s0 > 3, 4, 3 - s0).astype(np.uint8)
        keep += [s0, s1]
        p = arr[j]
        p.seq = s0.ctypes.data_as(C.POINTER(C.c_ubyte))
        p.rseq = s1.ctypes.data_as(C.POINTER(C.c_ubyte))
        p.len = p.full_len = r.size
        p.clip_len = r.size
        p.tid = -1
    return arr, keep
