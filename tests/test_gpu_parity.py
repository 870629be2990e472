"""GPU parity tests proper: every call goes through the C-ABI of libbwagpu.so (ctypes), and
is compared bit-for-bit with (a) the committed golden vectors the reference produced and
(b) the reference itself (oracle/_ref, prebuilt, travels with the snapshot) on fresh seeded
inputs.  Integer/byte work: the bar is bit-exact."""
import ctypes as C

import numpy as np
import pytest

import refload as R
from test_kernel_logic import CONFIGS, golden_case

pytestmark = pytest.mark.gpu
abi, api = R.abi, R.bwa.api


@pytest.fixture(scope="module")
def gpu_index(small_index):
    T, idx = small_index
    api.init()
    api.load_index(idx)
    yield T, idx
    api.destroy()


@pytest.mark.parametrize("name", CONFIGS)
def test_aln_flat_matches_golden(golden, gpu_index, name):
    reads, opt, want = golden_case(golden, name)
    got = api.aln_flat(reads.bases, reads.offs, opt)
    assert R.compare_aln(want, got, name) == []


def test_struct_api_matches_golden(golden, gpu_index):
    """bwa_gpu_cal_sa_reads_gap on bwa_seq_t[]: same fields the reference call fills
    (bwtaln.c:113,132), aln arrays libc-owned."""
    reads, opt, want = golden_case(golden, "ragged")
    seqs, keep = abi.make_seqs(reads)
    for s in seqs:  # poison the fields the call must reset
        s.sa, s.type, s.c1, s.c2, s.n_aln, s.max_entries = 77, 3, 5, 6, 99, -5
    api.cal_sa_reads_gap(seqs, opt)
    n_aln, max_entries, off, aln = want
    libc = C.CDLL(None)
    libc.free.argtypes = [C.c_void_p]
    for i, s in enumerate(seqs):
        assert s.n_aln == n_aln[i]
        assert (s.sa, s.type, s.c1, s.c2) == (0, 0, 0, 0)
        if s.len == 0:
            assert not s.aln
            continue
        assert bool(s.aln)  # bwt_match_gap always hands back a calloc'd array (bwtgap.c:115)
        if max_entries[i] != 0:
            assert s.max_entries == max_entries[i]
        else:
            assert s.max_entries == -5  # too many N: the reference leaves it untouched (bwtgap.c:120-123)
        if s.n_aln:
            buf = (C.c_char * (16 * s.n_aln)).from_address(C.addressof(s.aln.contents))
            assert bytes(buf) == aln[off[i]:off[i + 1]].tobytes()
        libc.free(C.cast(s.aln, C.c_void_p))


def test_cal_pac_pos_matches_golden(golden, gpu_index):
    out = api.cal_pac_pos(golden["sa_k"], golden["sa_which"])
    assert np.array_equal(out, golden["sa_out"])


def test_chunking_and_tiers_are_invisible(golden, gpu_index, monkeypatch):
    """Tiny chunks + a tiny tier-1 arena must not change a single byte."""
    reads, opt, want = golden_case(golden, "pe100")
    monkeypatch.setenv("BWAGPU_CHUNK", "97")
    monkeypatch.setenv("BWAGPU_T1_CAP", "48")
    monkeypatch.setenv("BWAGPU_POOL_MB", "1")         # 51 chunks: the optimistic pass runs dry
    monkeypatch.setenv("BWAGPU_HITS_PER_READ", "1")   # and so does the hit pool: grown, reads retried
    # the shared pool is sized on a lane's first use: a fresh context, so that the 1 MB really applies to every lane
    T, idx = gpu_index
    api.destroy()
    api.init()
    api.load_index(idx)
    try:
        got = api.aln_flat(reads.bases, reads.offs, opt)
        st = api.get_stats()
        assert st["x_chunks_used"] > 0      # searches spilled into the shared pool
        assert st["n_overflow_t2"] > 0      # reads left pass 0 (arena of 48 records)
        assert R.compare_aln(want, got, "chunked") == []
    finally:  # the module's other tests get a context with the default sizes back
        monkeypatch.undo()
        api.destroy()
        api.init()
        api.load_index(idx)


@pytest.mark.parametrize("name", CONFIGS)
def test_warp_pass_matches_golden(golden, gpu_index, monkeypatch, name):
    """A 2-record pass-0 arena sends (nearly) every read through k_search_warp (csrc/search_warp.cuh): one warp per read,
    the lanes draining the lowest bucket's top entries as independent chains.  Same bytes."""
    reads, opt, want = golden_case(golden, name)
    monkeypatch.setenv("BWAGPU_T1_CAP", "2")
    got = api.aln_flat(reads.bases, reads.offs, opt)
    st = api.get_stats()
    assert st["n_overflow_t2"] > reads.n // 2 and st["n_overflow_t3"] == 0   # they went through the warp pass and finished there
    assert R.compare_aln(want, got, "warp pass " + name) == []


@pytest.mark.parametrize("team", ["0", "1"])
@pytest.mark.parametrize("name", ["adna", "pe100", "nonstop_loggap"])
def test_warp_pass_both_forms(golden, gpu_index, monkeypatch, name, team):
    """BWAGPU_WARP_TEAM forces one warp per read (0) or the four warps of a block on one read (1); unset, the pass size decides."""
    reads, opt, want = golden_case(golden, name)
    monkeypatch.setenv("BWAGPU_T1_CAP", "2")
    monkeypatch.setenv("BWAGPU_WARP_TEAM", team)
    got = api.aln_flat(reads.bases, reads.offs, opt)
    assert api.get_stats()["n_overflow_t2"] > reads.n // 2
    assert R.compare_aln(want, got, f"warp pass form {team} {name}") == []


def test_thread_pass_still_available(golden, gpu_index, monkeypatch):
    """BWAGPU_WARP_PASS=0: pass 1 is the pooled thread-per-read kernel again (A/B switch)."""
    reads, opt, want = golden_case(golden, "adna")
    monkeypatch.setenv("BWAGPU_T1_CAP", "2")
    monkeypatch.setenv("BWAGPU_WARP_PASS", "0")
    got = api.aln_flat(reads.bases, reads.offs, opt)
    assert R.compare_aln(want, got, "thread pass") == []


@pytest.mark.skipif(not R.have_ref(), reason="oracle/_ref not present")
@pytest.mark.parametrize("cap", ["2", "2048"])
@pytest.mark.parametrize("label,length,optkw,simkw,n", [
    ("adna deep", (30, 50), dict(seed_len=1024, fnr=0.01, max_gapo=2), dict(adna=True, sub_rate=0.06), 20000),
    ("equal penalties: shared target bucket", 50, dict(s_mm=4, s_gapo=9, s_gape=4, max_gapo=2), dict(sub_rate=0.04), 5000),
    ("all penalties equal", 50, dict(s_mm=3, s_gapo=3, s_gape=3, max_gapo=2), dict(sub_rate=0.04), 5000),
    ("zero gap extension: serial rounds", 50, dict(s_mm=3, s_gapo=11, s_gape=0, max_gapo=1), dict(sub_rate=0.03), 3000),
    ("max_entries stop", (30, 50), dict(seed_len=1024, fnr=0.01, max_gapo=2, max_entries=3000), dict(adna=True, sub_rate=0.05), 10000),
    ("ragged, N", (15, 250), dict(max_gapo=2, max_gape=10), dict(n_rate=0.005), 10000),
])
def test_warp_pass_matches_reference_live(gpu_index, monkeypatch, cap, label, length, optkw, simkw, n):
    T, idx = gpu_index
    reads = R.bwa.simulate.simulate_reads(T, n, length, seed=777, **simkw)
    opt = abi.default_gap_opt(**optkw)
    want = R.ref_aln(R.RefIndex(idx), reads, opt, threads=8)
    monkeypatch.setenv("BWAGPU_T1_CAP", cap)
    got = api.aln_flat(reads.bases, reads.offs, opt)
    assert R.compare_aln(want, got, f"warp pass live, {label}, cap {cap}") == []


def test_stats_counters(golden, gpu_index):
    reads, opt, want = golden_case(golden, "se76")
    api.set_stats(True)
    try:
        got = api.aln_flat(reads.bases, reads.offs, opt)
        st = api.get_stats()
    finally:
        api.set_stats(False)
    assert R.compare_aln(want, got, "stats build") == []
    assert st["n_reads"] == reads.n and st["n_aln"] == want[3].size
    assert st["n_pops"] > reads.n and st["n_pushes"] >= st["n_pops"]
    assert st["occ_fetches_search"] > 0 and st["own_fetches_search"] > 0


def test_resident_path_matches_flat(golden, gpu_index):
    reads, opt, want = golden_case(golden, "se36")
    api.resident_stage(reads.bases, reads.offs, opt)
    ms1 = api.resident_run()
    ms2 = api.resident_run()  # idempotent: widths are recomputed, nothing carries over
    assert ms1 > 0 and ms2 > 0
    got = api.resident_fetch(reads.n)
    assert R.compare_aln(want, got, "resident") == []


@pytest.mark.skipif(not R.have_ref(), reason="oracle/_ref not present")
@pytest.mark.parametrize("length,optkw,simkw", [
    (36, {}, {}),
    (76, {}, {}),
    (100, {}, {}),
    ((30, 50), dict(seed_len=1024, fnr=0.01, max_gapo=2), dict(adna=True, sub_rate=0.01)),
    ((15, 250), dict(max_gapo=2, max_gape=10), dict(n_rate=0.005)),
    ((230, 400), {}, dict(sub_rate=0.01)),  # straddles 255 bases: longer reads skip pass 0 (its packed records hold 8-bit positions)
])
def test_aln_matches_reference_live(gpu_index, length, optkw, simkw):
    T, idx = gpu_index
    ridx = R.RefIndex(idx)
    reads = R.bwa.simulate.simulate_reads(T, 20000, length, seed=4242, **simkw)
    opt = abi.default_gap_opt(**optkw)
    want = R.ref_aln(ridx, reads, opt, threads=8)
    got = api.aln_flat(reads.bases, reads.offs, opt)
    assert R.compare_aln(want, got, f"live {length}") == []


@pytest.mark.skipif(not R.have_ref(), reason="oracle/_ref not present")
def test_cal_pac_pos_matches_reference_live(gpu_index):
    T, idx = gpu_index
    ridx = R.RefIndex(idx)
    rng = np.random.default_rng(8)
    k = rng.integers(0, idx.bwt[0].seq_len + 1, size=300000, dtype=np.uint32)
    which = rng.integers(0, 2, size=k.size, dtype=np.uint8)
    assert np.array_equal(api.cal_pac_pos(k, which), R.ref_sa(ridx, k, which))


def test_sa_of_every_row_is_a_permutation(gpu_index):
    """Size-independent property: bwt_sa over all rows 1..n of one strand is a permutation
    of 0..n-1 (row 0 is the sentinel row and yields -1 + steps, bwt.c:79-80)."""
    T, idx = gpu_index
    n = idx.bwt[0].seq_len
    k = np.arange(1, n + 1, dtype=np.uint32)
    for which in (1, 0):
        out = api.cal_pac_pos(k, np.full(k.size, which, dtype=np.uint8))
        assert np.array_equal(np.sort(out), np.arange(n, dtype=np.uint32))


def test_exact_reads_are_found_at_their_origin(gpu_index):
    """Size-independent property used at full size in bench.py --check: an error-free read
    cut from the genome gets a 0-difference hit whose SA interval contains its origin."""
    T, idx = gpu_index
    reads = R.bwa.simulate.simulate_reads(T, 5000, 50, seed=77, sub_rate=0.0, indel_frac=0.0, n_rate=0.0, junk_frac=0.0)
    opt = abi.default_gap_opt()
    n_aln, _, off, aln = api.aln_flat(reads.bases, reads.offs, opt)
    assert (n_aln >= 1).all()
    first = aln[off[:-1]]
    assert (first["score"] == 0).all()
    # expand the first hit of each read to coordinates and look for the simulated origin
    ks, which, owner = [], [], []
    for i in range(reads.n):
        a = first[i]
        strand = (a["info"] >> 24) & 1
        w = min(int(a["l"] - a["k"] + 1), 64)
        ks.append(np.arange(a["k"], a["k"] + w, dtype=np.uint32))
        which.append(np.full(w, strand, dtype=np.uint8))
        owner.append(np.full(w, i))
    ks, which, owner = np.concatenate(ks), np.concatenate(which), np.concatenate(owner)
    sa = api.cal_pac_pos(ks, which)
    n = idx.bwt[0].seq_len
    pos = np.where(which == 1, sa, n - (sa + 50)).astype(np.int64)  # bwase.c:144-153
    hit = np.zeros(reads.n, dtype=bool)
    np.logical_or.at(hit, owner, pos == reads.pos[owner])
    full = (first["l"] - first["k"] + 1) <= 64
    assert hit[full].all()


def sw_jobs_from(begs, reglens, queries, q_off):
    return [(int(begs[i]), int(reglens[i]), queries[q_off[i]:q_off[i + 1]]) for i in range(len(begs))]


def test_mate_sw_matches_golden(golden, gpu_index):
    """K5: score, start and end cells of aln_local_core on pac windows (bwape.c:447-456)."""
    jobs = sw_jobs_from(golden["swp_beg"], golden["swp_reglen"], golden["swp_queries"], golden["swp_q_off"])
    got = np.array(api.mate_sw(jobs), dtype=np.int32)
    want = golden["swp_out"]
    bad = np.nonzero((got != want).any(1))[0]
    assert bad.size == 0, (bad[:5], got[bad[:5]], want[bad[:5]])


@pytest.mark.skipif(not R.have_ref(), reason="oracle/_ref not present")
def test_mate_sw_matches_reference_live(gpu_index):
    T, idx = gpu_index
    # wide windows (several 512-column sweeps), long and short reads, window clipped at the genome end
    refs, ro, qs, qo, begs = R.make_sw_jobs(T, 6000, seed=1234, ref_n=False, with_beg=True, read_len=(20, 260), win=(20, 1400))
    reglens = np.diff(ro).astype(np.int32)
    want = R.ref_sw_batch(refs, ro, qs, qo)
    got = np.array(api.mate_sw(sw_jobs_from(begs, reglens, qs, qo)), dtype=np.int32)
    bad = np.nonzero((got != want).any(1))[0]
    assert bad.size == 0, (bad[:5], got[bad[:5]], want[bad[:5]])
    # a window running past l_pac is clipped like bwa_sw_core's copy loop
    n = idx.l_pac
    q = T[n - 60:n - 10].copy()
    got = api.mate_sw([(n - 100, 400, q)])
    win = T[n - 100:n]
    want = R.ref_sw_batch(win, np.array([0, win.size]), q, np.array([0, q.size]))
    assert list(got[0]) == list(want[0])


def test_mate_sw_empty_inputs(gpu_index):
    T, idx = gpu_index
    assert api.mate_sw([]) == []
    got = api.mate_sw([(10, 0, T[:30]), (10, 50, T[:0])])
    assert got[0][0] == -1 and got[1][0] == -1  # stdaln.c:559


def test_multi_device_in_process(golden, small_index):
    """All visible GPUs inside one process (bwa_gpu_init with several ids): the index is replicated,
    reads are split into per-device ranges, results come back in read order."""
    import torch
    n_dev = torch.cuda.device_count()
    if n_dev < 2:
        pytest.skip("needs >= 2 GPUs")
    T, idx = small_index
    api.destroy()
    api.init(list(range(n_dev)))
    try:
        api.load_index(idx)
        for name in ("pe100", "ragged"):
            reads, opt, want = golden_case(golden, name)
            got = api.aln_flat(reads.bases, reads.offs, opt)
            assert R.compare_aln(want, got, f"{n_dev} devices / {name}") == []
        assert api.get_stats()["n_devices"] == n_dev
        out = api.cal_pac_pos(golden["sa_k"], golden["sa_which"])
        assert np.array_equal(out, golden["sa_out"])
    finally:
        api.destroy()


def _same_path(got, want):
    return tuple(got[:5]) == tuple(want[:5]) and np.array_equal(got[5], want[5])


@pytest.mark.skipif(not R.have_ref(), reason="oracle/_ref not present")
def test_mate_sw_path_matches_reference_live(gpu_index):
    """K5 + K6: the whole aln_local_core call of bwa_sw_core -- score, path end points and CIGAR."""
    T, idx = gpu_index
    refs, ro, qs, qo, begs = R.make_sw_jobs(T, 2500, seed=321, ref_n=False, with_beg=True, read_len=(20, 200), win=(20, 900))
    jobs = sw_jobs_from(begs, np.diff(ro).astype(np.int32), qs, qo)
    got = api.mate_sw_path(jobs)
    bad = []
    for i in range(len(jobs)):
        want = R.ref_sw_path(refs[ro[i]:ro[i + 1]], qs[qo[i]:qo[i + 1]])
        if want[0] < 1:  # nothing aligned: the reference returns before making a path
            if not (got[i][0] == want[0] and got[i][5].size == 0):
                bad.append(i)
        elif not _same_path(got[i], want):
            bad.append(i)
    assert not bad, (len(bad), bad[:5], got[bad[0]], R.ref_sw_path(refs[ro[bad[0]]:ro[bad[0] + 1]], qs[qo[bad[0]]:qo[bad[0] + 1]]))


@pytest.mark.skipif(not R.have_ref(), reason="oracle/_ref not present")
@pytest.mark.parametrize("gap_end,band", [(5, 50), (-1, 50), (5, 7), (-1, 2)])
def test_global_align_matches_reference_live(gpu_index, gap_end, band):
    """K6 alone: aln_global_core as refine_gapped_core calls it (gap_end 5, band 50) and with other settings."""
    T, idx = gpu_index
    rng = np.random.default_rng(band)
    refs, ro, qs, qo, begs = R.make_sw_jobs(T, 1500, seed=100 + band, ref_n=False, with_beg=True, read_len=(20, 150), win=(160, 161))
    jobs, wants = [], []
    for i in range(1500):
        q = qs[qo[i]:qo[i + 1]]
        rl = int(q.size + rng.integers(0, 9))  # len + |ext| like refine_gapped_core's window
        beg = int(begs[i] + rng.integers(0, 20))
        jobs.append((beg, rl, q))
        wants.append(R.ref_global(T[beg:beg + rl], q, gap_end, band))
    got = api.global_align(jobs, gap_end, band)
    bad = [i for i in range(1500) if not _same_path(got[i], wants[i])]
    assert not bad, (len(bad), bad[:5], got[bad[0]], wants[bad[0]])


@pytest.mark.skipif(not R.have_ref(), reason="oracle/_ref not present")
def test_global_align_both_kernels(gpu_index):
    """K6's two forms in one call: jobs that fit the warp-per-job kernel's shared memory (csrc/sw.cuh k_global_warp: window
    <= 256 columns, band cells <= 16 K) next to jobs that do not and go to the thread-per-job kernel, in random order."""
    T, idx = gpu_index
    rng = np.random.default_rng(77)
    jobs, wants = [], []
    for i in range(600):
        ql = int(rng.integers(20, 120)) if i % 3 else int(rng.integers(240, 420))
        beg = int(rng.integers(0, idx.l_pac - 600))
        q = T[beg + 3:beg + 3 + ql].copy()
        for _ in range(int(rng.integers(0, 5))):
            q[int(rng.integers(0, ql))] = rng.integers(0, 4)
        if i % 4 == 0 and ql > 30:  # a deletion in the read
            cut = int(rng.integers(5, ql - 10))
            q = np.concatenate([q[:cut], q[cut + int(rng.integers(1, 4)):]])
        rl = int(q.size + rng.integers(0, 9))
        jobs.append((beg, rl, q))
        wants.append(R.ref_global(T[beg:beg + rl], q, 5, 50))
    got = api.global_align(jobs, 5, 50)
    bad = [i for i in range(len(jobs)) if not _same_path(got[i], wants[i])]
    assert not bad, (len(bad), bad[:5], got[bad[0]], wants[bad[0]])


def test_mate_sw_path_and_global_align_match_golden(golden, gpu_index):
    """K6 against the committed reference-made vectors (no oracle/_ref needed)."""
    jobs = sw_jobs_from(golden["swp_beg"], golden["swp_reglen"], golden["swp_queries"], golden["swp_q_off"])
    got = api.mate_sw_path(jobs)
    want, wc, wo = golden["swpath_out"], golden["swpath_cigar"], golden["swpath_cigar_off"]
    for i in range(len(jobs)):
        if want[i, 0] < 1:
            assert got[i][0] == want[i, 0] and got[i][5].size == 0, i
        else:
            assert tuple(got[i][:5]) == tuple(want[i]) and np.array_equal(got[i][5], wc[wo[i]:wo[i + 1]]), i
    par = golden["glob_par"]
    qo, co = golden["glob_q_off"], golden["glob_cigar_off"]
    for ge, band in sorted({(int(a), int(b)) for a, b in par}):
        sel = [i for i in range(par.shape[0]) if (int(par[i, 0]), int(par[i, 1])) == (ge, band)]
        jobs = [(int(golden["glob_beg"][i]), int(golden["glob_reglen"][i]), golden["glob_queries"][qo[i]:qo[i + 1]]) for i in sel]
        got = api.global_align(jobs, ge, band)
        for g, i in zip(got, sel):
            assert tuple(g[:5]) == tuple(golden["glob_out"][i]) and np.array_equal(g[5], golden["glob_cigar"][co[i]:co[i + 1]]), (i, ge, band)


def test_argument_errors_are_loud(gpu_index):
    """Unsupported inputs fail with a message instead of computing something else."""
    T, idx = gpu_index
    opt = abi.default_gap_opt()
    long_read = np.zeros(40000, dtype=np.uint8)
    with pytest.raises(api.BwaGpuError, match="read length"):
        api.aln_flat(long_read, np.array([0, long_read.size], dtype=np.int64), opt)
    big = abi.default_gap_opt(fnr=-1.0, max_diff=300)
    with pytest.raises(api.BwaGpuError, match="not supported|exceeds"):
        api.aln_flat(T[:50].copy(), np.array([0, 50], dtype=np.int64), big)
    wide = abi.default_gap_opt(s_mm=40, s_gapo=90, s_gape=30)
    with pytest.raises(api.BwaGpuError, match="buckets"):
        api.aln_flat(T[:50].copy(), np.array([0, 50], dtype=np.int64), wide)
    with pytest.raises(api.BwaGpuError, match="out of range"):
        api.cal_pac_pos(np.array([idx.bwt[0].seq_len + 5], dtype=np.uint32), np.array([1], dtype=np.uint8))
    # an empty batch is fine
    n_aln, me, off, aln = api.aln_flat(np.zeros(0, np.uint8), np.zeros(1, np.int64), opt)
    assert n_aln.size == 0 and off.tolist() == [0] and aln.size == 0
    assert api.cal_pac_pos(np.zeros(0, np.uint32), np.zeros(0, np.uint8)).size == 0


def test_many_buckets_option_set(gpu_index):
    """More than 128 score buckets (4 mask words in k_search): -M 5 -O 20 -E 8 with -n 6."""
    T, idx = gpu_index
    if not R.have_ref():
        pytest.skip("oracle/_ref not present")
    reads = R.bwa.simulate.simulate_reads(T, 3000, 60, seed=66)
    opt = abi.default_gap_opt(s_mm=5, s_gapo=20, s_gape=8, fnr=-1.0, max_diff=6, max_gapo=2)
    want = R.ref_aln(R.RefIndex(idx), reads, opt, threads=8)
    got = api.aln_flat(reads.bases, reads.offs, opt)
    assert R.compare_aln(want, got, "many buckets") == []
