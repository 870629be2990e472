"""Pins oracle/bwa_oracle.c (this repo's CPU restatement, the checker) against golden vectors
the reference itself produced, and against the reference live when oracle/_ref is present."""
import numpy as np
import pytest

import refload as R
from test_kernel_logic import CONFIGS, golden_case


@pytest.fixture(scope="module")
def oidx(small_index):
    T, idx = small_index
    return R.orc_index(idx), idx


@pytest.mark.parametrize("name", CONFIGS)
def test_oracle_aln_matches_golden(golden, oidx, name):
    reads, opt, want = golden_case(golden, name)
    got = R.orc_aln(oidx[0], reads, opt)
    # the oracle leaves max_entries at 0 where the reference leaves the field untouched
    assert R.compare_aln(want, got, name) == []


def test_oracle_sa_matches_golden(golden, oidx):
    O = R.orc()
    k, which, want = golden["sa_k"], golden["sa_which"], golden["sa_out"]
    for i in range(0, k.size, 7):
        s = 0 if which[i] else 1
        assert O.orc_sa(oidx[0][s], int(k[i])) == int(want[i])


def test_oracle_maxdiff_matches_golden(golden):
    O = R.orc()
    for name, fnr in (("maxdiff_004", 0.04), ("maxdiff_001", 0.01)):
        want = golden[name]
        got = [O.orc_cal_maxdiff(l, 0.02, float(np.float32(fnr))) for l in range(want.size)]
        assert got == list(want)


def test_oracle_occ_bruteforce(small_index, oidx):
    """orc_occ against a direct count on the BWT rebuilt from the suffix array definition."""
    import torch
    T, idx = small_index
    n = 3000
    t = torch.from_numpy(T[:n].copy())
    small = R.bwa.index.build_bwt(t)
    o = R.orc_index(R.bwa.index.FMIndex(bwt=[small, small], pac=np.zeros(1, np.uint8), l_pac=n))
    sa = R.bwa.index.suffix_array(t).numpy()
    full = np.concatenate([[n], sa])
    bw = np.where(full > 0, T[:n][full - 1], 9)  # 9 marks the '$' row
    O = R.orc()
    rng = np.random.default_rng(0)
    for k in list(rng.integers(0, n + 1, size=300)) + [0, n, small.primary, small.primary - 1]:
        for c in range(4):
            assert O.orc_occ(o[0], int(k), c) == int((bw[: k + 1] == c).sum())


def test_oracle_sw_matches_golden(golden):
    got = R.orc_sw_batch(golden["sw_refs"], golden["sw_ref_off"], golden["sw_queries"], golden["sw_q_off"])
    assert np.array_equal(got, golden["sw_out"])


def test_oracle_sw_edge_cases():
    O = R.orc()
    import ctypes as C
    res = (C.c_int * 4)()
    a = np.array([0, 1, 2, 3, 0, 1, 2, 3], dtype=np.uint8)
    assert O.orc_sw_local(a.ctypes.data, 0, a.ctypes.data, 8, res) == -1  # empty input (stdaln.c:559)
    n4 = np.full(8, 4, dtype=np.uint8)
    assert O.orc_sw_local(n4.ctypes.data, 8, n4.ctypes.data, 8, res) == 0 and list(res) == [0, 0, 0, 0]
    assert O.orc_sw_local(a.ctypes.data, 8, a.ctypes.data, 8, res) == 88 and list(res) == [1, 1, 8, 8]


@pytest.mark.skipif(not R.have_ref(), reason="oracle/_ref not built")
def test_oracle_sw_matches_reference_live(small_index):
    T, _ = small_index
    refs, ro, qs, qo = R.make_sw_jobs(T, 1500, seed=99, read_len=(20, 150), win=(40, 500))
    assert np.array_equal(R.orc_sw_batch(refs, ro, qs, qo), R.ref_sw_batch(refs, ro, qs, qo))


@pytest.mark.skipif(not R.have_ref(), reason="oracle/_ref not built")
def test_oracle_aln_matches_reference_live(small_index, oidx):
    T, idx = small_index
    reads = R.bwa.simulate.simulate_reads(T, 600, (18, 110), seed=31, n_rate=0.01)
    opt = R.abi.default_gap_opt(max_gapo=2)
    want = R.ref_aln(R.RefIndex(idx), reads, opt, threads=4)
    assert R.compare_aln(want, R.orc_aln(oidx[0], reads, opt), "live") == []


def _orc_path(fn, *args):
    import ctypes as C
    n = args[1] + args[3] + 4
    path = np.zeros(3 * n, dtype=np.int32)
    plen = C.c_int()
    score = fn(*args, path, plen)
    p = path[:3 * plen.value].reshape(-1, 3)
    if plen.value == 0:
        return (score, 0, 0, 0, 0, np.empty(0, np.uint16))
    return (score, int(p[-1, 0]), int(p[-1, 1]), int(p[0, 0]), int(p[0, 1]), R.path_to_cigar(p))


def test_oracle_local_path_matches_golden(golden):
    """orc_sw_local_path = the whole aln_local_core call incl. the banded global third pass."""
    import ctypes as C
    O = R.orc()
    O.orc_sw_local_path.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_int)]
    T = golden["genome"]
    begs, reglen, qs, qo = golden["swp_beg"], golden["swp_reglen"], golden["swp_queries"], golden["swp_q_off"]
    want, wc, wo = golden["swpath_out"], golden["swpath_cigar"], golden["swpath_cigar_off"]
    for i in range(0, begs.size, 3):
        r = np.ascontiguousarray(T[begs[i]:begs[i] + reglen[i]]); q = np.ascontiguousarray(qs[qo[i]:qo[i + 1]])
        got = _orc_path(lambda a, b, c, d, path, plen: O.orc_sw_local_path(a, b, c, d, path.ctypes.data, C.byref(plen)),
                        r.ctypes.data, r.size, q.ctypes.data, q.size)
        if want[i, 0] < 1:
            assert got[0] == want[i, 0]
        else:
            assert tuple(got[:5]) == tuple(want[i]) and np.array_equal(got[5], wc[wo[i]:wo[i + 1]]), i


def test_oracle_global_matches_golden(golden):
    import ctypes as C
    O = R.orc()
    O.orc_global.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_int)]
    T = golden["genome"]
    for i in range(golden["glob_beg"].size):
        b, rl = int(golden["glob_beg"][i]), int(golden["glob_reglen"][i])
        r = np.ascontiguousarray(T[b:b + rl]); q = np.ascontiguousarray(golden["glob_queries"][golden["glob_q_off"][i]:golden["glob_q_off"][i + 1]])
        ge, band = (int(x) for x in golden["glob_par"][i])
        got = _orc_path(lambda a, b_, c, d, path, plen: O.orc_global(a, b_, c, d, ge, band, path.ctypes.data, C.byref(plen)),
                        r.ctypes.data, r.size, q.ctypes.data, q.size)
        co = golden["glob_cigar_off"]
        assert tuple(got[:5]) == tuple(golden["glob_out"][i]) and np.array_equal(got[5], golden["glob_cigar"][co[i]:co[i + 1]]), i
