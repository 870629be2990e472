"""CPU checks of the kernels' LOGIC: the bodies of csrc/kernels.cuh compiled for the host
(tests/host_emu) against (a) the committed golden vectors made by the reference and (b) the
reference itself when oracle/_ref is present.  These do not replace the GPU parity tests;
they catch traversal-order and layout mistakes without a device."""
import ctypes as C

import numpy as np
import pytest

import refload as R

abi = R.abi
CONFIGS = ["se36", "se76", "pe100", "adna", "ragged", "fixed_n3", "nonstop_loggap", "nogape"]


def opt_from_words(words):
    o = abi.gap_opt_t()
    C.memmove(C.addressof(o), words.tobytes(), 64)
    return o


def golden_case(golden, name):
    reads = R.bwa.simulate.Reads(golden[f"{name}_bases"], golden[f"{name}_offs"], None, None)
    opt = opt_from_words(golden[f"{name}_opt"])
    n_aln = golden[f"{name}_n_aln"]
    off = np.zeros(n_aln.size + 1, dtype=np.int64)
    off[1:] = np.cumsum(n_aln)
    aln = np.ascontiguousarray(golden[f"{name}_aln"]).view(abi.ALN_DTYPE).reshape(-1)
    return reads, opt, (n_aln, golden[f"{name}_max_entries"], off, aln)


@pytest.fixture(scope="module")
def emu_index(small_index):
    T, idx = small_index
    ridx = R.RefIndex(idx)
    h = R.emu().emu_load_index(ridx.arr)
    yield h, ridx
    R.emu().emu_free_index(h)


@pytest.mark.parametrize("name", CONFIGS)
def test_search_logic_matches_golden(golden, emu_index, name):
    h, _ = emu_index
    reads, opt, want = golden_case(golden, name)
    got = R.emu_aln(h, reads, opt)
    assert R.compare_aln(want, got, name) == []


def test_search_logic_pool_chunks(golden, emu_index):
    """Tiny private arenas: almost every search spills into chunks of the shared pool."""
    h, _ = emu_index
    for name in ("se76", "adna"):
        reads, opt, want = golden_case(golden, name)
        got = R.emu_aln(h, reads, opt, cap1=8, aln_cap1=64, n_slots=2, pool_chunks=4096)
        assert got[4][4] == 0  # nothing needed the guaranteed pass
        assert R.compare_aln(want, got, name) == []


def test_search_logic_guaranteed_pass(golden, emu_index):
    """A pool that runs dry pushes reads into the guaranteed pass."""
    h, _ = emu_index
    reads, opt, want = golden_case(golden, "se76")
    got = R.emu_aln(h, reads, opt, cap1=32, aln_cap1=1, n_slots=2, pool_chunks=0)
    assert got[4][4] > 10  # reads retried in pass 1
    assert R.compare_aln(want, got, "guaranteed") == []


def test_search_logic_pop_cap(golden, emu_index, monkeypatch):
    """Pass 0 hands a read that pops more than Batch::pop_cap nodes to the next pass (bounded tail of the thread-per-read kernel);
    whatever it had found or edited (gap_shadow) by then is dropped and the retry starts from pristine widths."""
    h, _ = emu_index
    monkeypatch.setenv("EMU_POP_CAP", "40")
    for name in ("se76", "adna"):
        reads, opt, want = golden_case(golden, name)
        got = R.emu_aln(h, reads, opt, cap1=2048, aln_cap1=64, n_slots=2, pool_chunks=4096)
        assert R.compare_aln(want, got, name) == []


def test_search_logic_retry_after_hits(golden, emu_index):
    """A read that fails AFTER its first hit has had its widths edited by gap_shadow;
    the retry must start from pristine widths."""
    h, _ = emu_index
    reads, opt, want = golden_case(golden, "adna")
    got = R.emu_aln(h, reads, opt, cap1=100, aln_cap1=64, n_slots=2, pool_chunks=0)
    assert got[4][4] > 20
    assert R.compare_aln(want, got, "retry") == []


def test_sa_logic_matches_golden(golden, emu_index):
    h, _ = emu_index
    k, which = golden["sa_k"], golden["sa_which"]
    out = np.empty(k.size, dtype=np.uint32)
    R.emu().emu_sa(h, k.size, k.ctypes.data, which.ctypes.data, out.ctypes.data)
    assert np.array_equal(out, golden["sa_out"])


@pytest.mark.skipif(not R.have_ref(), reason="oracle/_ref not built")
def test_search_logic_matches_reference_live(small_index, emu_index):
    T, _ = small_index
    h, ridx = emu_index
    reads = R.bwa.simulate.simulate_reads(T, 800, (20, 120), seed=123, n_rate=0.01)
    opt = abi.default_gap_opt(max_gapo=2, max_gape=8)
    want = R.ref_aln(ridx, reads, opt, threads=4)
    got = R.emu_aln(h, reads, opt)
    assert R.compare_aln(want, got, "live") == []
    # reads on both sides of 255 bases: pass 0's records pack positions into 8 bits, longer reads start in the next pass
    reads = R.bwa.simulate.simulate_reads(T, 60, (240, 300), seed=99, sub_rate=0.01)
    assert (np.diff(reads.offs) > 255).sum() > 10
    opt = abi.default_gap_opt()
    assert R.compare_aln(R.ref_aln(ridx, reads, opt, threads=4), R.emu_aln(h, reads, opt), "live long") == []


# ---------------------------------------------------------------- k_search_warp (csrc/search_warp.cuh) on an emulated 32-lane warp
def _prefix(reads, want, n):
    r = R.bwa.simulate.Reads(reads.bases[:reads.offs[n]], reads.offs[:n + 1], None, None)
    na, me, off, aln = want
    return r, (na[:n], me[:n], off[:n + 1], aln[:off[n]])


@pytest.fixture(scope="module")
def wemu_index(small_index):
    T, idx = small_index
    ridx = R.RefIndex(idx)
    h = R.wemu().wemu_load_index(ridx.arr)
    yield h, ridx
    R.wemu().wemu_free_index(h)


@pytest.mark.parametrize("name", CONFIGS)
def test_warp_search_logic_matches_golden(golden, wemu_index, name):
    """The warp-per-read kernel body, lanes as coroutines with real shuffles / ballots (tests/host_emu/warp_emu.cpp): chains per
    lane, count / scan / store passes, rollback after a hit, chunked buckets -- same bytes as the reference."""
    h, _ = wemu_index
    reads, opt, want = golden_case(golden, name)
    r, w = _prefix(reads, want, 60 if name == "nonstop_loggap" else 150)
    got = R.wemu_aln(h, r, opt)
    assert got[4] == 0
    assert R.compare_aln(w, got, "warp " + name) == []


@pytest.mark.parametrize("name", ["se76", "adna", "ragged", "nogape"])
def test_team_search_logic_matches_golden(golden, small_index, name):
    """The same kernel body with the four warps of a block sharing one read (rounds of up to 128 chains, block-wide scans)."""
    T, idx = small_index
    ridx = R.RefIndex(idx)
    h = R.wemu(4).wemu_load_index(ridx.arr)
    try:
        reads, opt, want = golden_case(golden, name)
        r, w = _prefix(reads, want, 80)
        got = R.wemu_aln(h, r, opt, team=4)
        assert got[4] == 0
        assert R.compare_aln(w, got, "team " + name) == []
    finally:
        R.wemu(4).wemu_free_index(h)


def test_warp_search_logic_pool_dry(golden, wemu_index):
    """A chunk pool too small for every read: the reads that find it dry are flagged (n_aln = -1, retried by the guaranteed pass
    in the library); the others are still exact."""
    h, _ = wemu_index
    reads, opt, want = golden_case(golden, "adna")
    r, w = _prefix(reads, want, 150)
    got = R.wemu_aln(h, r, opt, pool_chunks=12)
    assert 0 < got[4] < 150 and int((got[0] < 0).sum()) == got[4]
    ok = got[0] >= 0
    assert np.array_equal(got[0][ok], w[0][ok]) and np.array_equal(got[1][ok], w[1][ok])


@pytest.mark.skipif(not R.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("optkw,simkw", [
    (dict(seed_len=1024, fnr=0.01, max_gapo=2), dict(adna=True, sub_rate=0.06)),   # deep searches
    (dict(s_mm=4, s_gapo=9, s_gape=4, max_gapo=2), dict(sub_rate=0.04)),           # two penalties equal: shared target bucket
    (dict(s_mm=3, s_gapo=11, s_gape=0, max_gapo=1), dict(sub_rate=0.03)),          # a zero penalty: one-lane rounds
    (dict(seed_len=1024, fnr=0.01, max_gapo=2, max_entries=3000), dict(adna=True, sub_rate=0.05)),  # the max_entries stop
])
def test_warp_search_logic_matches_reference_live(small_index, wemu_index, optkw, simkw):
    T, _ = small_index
    h, ridx = wemu_index
    reads = R.bwa.simulate.simulate_reads(T, 120, (30, 50), seed=321, **simkw)
    opt = abi.default_gap_opt(**optkw)
    want = R.ref_aln(ridx, reads, opt, threads=4)
    got = R.wemu_aln(h, reads, opt)
    assert got[4] == 0
    assert R.compare_aln(want, got, "warp live") == []


@pytest.mark.skipif(not R.have_ref(), reason="oracle/_ref not built")
def test_option_fuzz_matches_reference_live(small_index, emu_index, wemu_index):
    """Random option sets (penalties, gap limits, seed, mode bits, max_top2, max_entries, max_diff given or derived) and read shapes:
    both kernel bodies against the live reference.  scripts/fuzz_kernel_logic.py runs the same loop for as long as one likes."""
    T, _ = small_index
    he, ridx = emu_index
    hw, _ = wemu_index
    rng = np.random.default_rng(2024)
    for it in range(12):
        kw = dict(s_mm=int(rng.integers(1, 6)), s_gapo=int(rng.integers(1, 14)), s_gape=int(rng.integers(1, 7)),
                  max_gapo=int(rng.integers(0, 3)), max_gape=int(rng.integers(0, 8)), indel_end_skip=int(rng.integers(0, 7)),
                  max_del_occ=int(rng.integers(1, 20)), seed_len=int(rng.choice([8, 16, 32, 1024])), max_seed_diff=int(rng.integers(0, 3)),
                  max_top2=int(rng.choice([0, 1, 3, 30])), max_entries=int(rng.choice([200, 3000])),
                  mode=int(rng.choice([0x01, 0x00, 0x05, 0x11, 0x15, 0x04, 0x10])) | 0x02)
        if rng.random() < 0.5:
            kw["fnr"] = float(rng.choice([0.04, 0.01, 0.1]))
        else:
            kw["fnr"], kw["max_diff"] = -1.0, int(rng.integers(0, 4))
        opt = abi.default_gap_opt(**kw)
        lo = int(rng.integers(8, 50))
        reads = R.bwa.simulate.simulate_reads(T, 40, (lo, lo + int(rng.integers(0, 40))), seed=int(rng.integers(1, 1 << 30)),
                                              sub_rate=float(rng.choice([0.0, 0.02, 0.06])), n_rate=float(rng.choice([0.0, 0.01, 0.05])))
        want = R.ref_aln(ridx, reads, opt, threads=4)
        got_w = R.wemu_aln(hw, reads, opt)
        assert got_w[4] == 0 and R.compare_aln(want, got_w, f"fuzz warp {kw}") == []
        assert R.compare_aln(want, R.emu_aln(he, reads, opt), f"fuzz thread {kw}") == []


@pytest.mark.skipif(not R.have_ref(), reason="oracle/_ref not present")
def test_zero_penalty_options_match_reference_live(small_index, emu_index):
    """One penalty at a time set to 0: a popped child's own pushes then land in the bucket it came from, on top of the group
    record it was popped from -- the case in which k_search's pop cache (the group record kept in shared memory while it has
    children left) has to write a displaced group's mask back to the arena."""
    T, _ = small_index
    he, ridx = emu_index
    rng = np.random.default_rng(11)
    for it in range(9):
        z = it % 3
        kw = dict(s_mm=0 if z == 0 else int(rng.integers(1, 4)), s_gapo=0 if z == 1 else int(rng.integers(1, 8)),
                  s_gape=0 if z == 2 else int(rng.integers(1, 4)), max_gapo=int(rng.integers(1, 3)), max_gape=int(rng.integers(1, 6)),
                  max_entries=3000, seed_len=int(rng.choice([16, 32, 1024])), max_top2=int(rng.choice([1, 30])),
                  mode=int(rng.choice([0x01, 0x11, 0x00])) | 0x02, fnr=-1.0, max_diff=int(rng.integers(1, 4)))
        opt = abi.default_gap_opt(**kw)
        reads = R.bwa.simulate.simulate_reads(T, 60, (20, 45), seed=int(rng.integers(1, 1 << 30)), sub_rate=0.05, n_rate=0.01)
        want = R.ref_aln(ridx, reads, opt, threads=4)
        assert R.compare_aln(want, R.emu_aln(he, reads, opt), f"zero penalty {kw}") == []


@pytest.mark.skipif(not R.have_ref(), reason="oracle/_ref not present")
def test_library_maxdiff_matches_reference_for_long_reads():
    """csrc/hostprep.h cal_maxdiff against the reference's bwa_cal_maxdiff (bwtaln.c:37-49) for every length 1..1000: from
    ~350 bp on the reference's `int x *= k` wraps around, and the restatement must wrap to the same bits (it multiplies in
    uint32_t instead of relying on signed overflow)."""
    E = R.emu()
    E.emu_cal_maxdiff.argtypes = [C.c_int, C.c_double, C.c_double]
    L = R.ref()[0]
    for fnr in (0.04, 0.01, 0.001, 0.2):
        f = float(np.float32(fnr))  # gap_opt_t.fnr is a float
        got = [E.emu_cal_maxdiff(l, 0.02, f) for l in range(1, 1001)]
        want = [L.bwa_cal_maxdiff(l, 0.02, f) for l in range(1, 1001)]
        assert got == want, (fnr, [(l + 1, a, b) for l, (a, b) in enumerate(zip(got, want)) if a != b][:5])


@pytest.mark.skipif(not R.have_ref(), reason="oracle/_ref not present")
@pytest.mark.parametrize("seed,read_len,win", [(1, (20, 150), (20, 500)), (2, (20, 260), (20, 1400)), (3, (1, 40), (1, 70))])
def test_sw_forward_cell_matches_reference(seed, read_len, win):
    """K5 pass 1 as the kernel computes it (csrc/sw_cell.h: G = H - qr state, F updated unconditionally, DPX-shaped max
    chains; csrc/sw.cuh sw_sweep: strips of C columns with padding, first-maximum keys), on the CPU around the kernel's own
    cell function (tests/host_emu/sw_emu.cpp), against aln_local_core's score and end cell (stdaln.c:608-627): windows of
    one to three 512-column sweeps, reads with N, indels, and reads that have nothing to do with their window."""
    import os
    import subprocess
    src = os.path.join(R.EMU_DIR, "sw_emu.cpp")
    so = os.path.join(R.EMU_DIR, "libsw_emu.so")
    cell = os.path.join(R.ROOT, "network-aware-bwa_b200", "csrc", "sw_cell.h")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(cell)):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", so, src], check=True)
    E = C.CDLL(so)
    E.sw_emu_pass1.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    T = R.bwa.simulate.make_genome(300000, seed=21, repeat_frac=0.05)
    n = 2000
    refs, ro, qs, qo = R.make_sw_jobs(T, n, seed=seed, ref_n=False, read_len=read_len, win=win)
    want = R.ref_sw_batch(refs, ro, qs, qo)
    out = (C.c_int * 3)()
    bad = []
    for i in range(n):
        r = np.ascontiguousarray(refs[ro[i]:ro[i + 1]])
        q = np.ascontiguousarray(qs[qo[i]:qo[i + 1]])
        E.sw_emu_pass1(r.ctypes.data, r.size, q.ctypes.data, q.size, out)
        w = want[i]
        ok = out[0] == w[0] if w[0] < 1 else (out[0], out[1], out[2]) == (w[0], w[3], w[4])
        if not ok:
            bad.append((i, list(out), list(w)))
    assert not bad, bad[:5]
