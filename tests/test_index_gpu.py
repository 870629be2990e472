"""The library's own index builder (csrc/indexbuild.cu, bwa_gpu_index_build) against the reference's `bwa index`
(live, byte for byte) and against the torch harness builder; and BASELINE.json configs[3] (C4): a genome of more than
3.0e9 bases -- SA values above 2^31, `k >= primary` shifts and `seq_len - (sa + len)` wraps in the upper half of the
32-bit range -- searched on the device and compared with the live reference on the same index."""
import os
import subprocess

import numpy as np
import pytest

import refload as R

pytestmark = pytest.mark.gpu

ix = R.bwa.index
sim = R.bwa.simulate
api = R.bwa.api
abi = R.abi


def same_index(a, b):
    for s in range(2):
        assert a.bwt[s].primary == b.bwt[s].primary, s
        assert np.array_equal(a.bwt[s].L2, b.bwt[s].L2), s
        assert a.bwt[s].seq_len == b.bwt[s].seq_len
        assert np.array_equal(a.bwt[s].bwt, b.bwt[s].bwt), s
        assert np.array_equal(a.bwt[s].sa, b.bwt[s].sa), s


@pytest.mark.parametrize("n", [1, 2, 15, 16, 17, 21, 22, 127, 128, 129, 2048, 4099, 100003])
def test_native_builder_equals_harness_builder(n):
    T = np.random.default_rng(n).integers(0, 4, size=n, dtype=np.uint8)
    same_index(ix.build_index_native(T), ix.build_index(T, device="cpu"))


@pytest.mark.parametrize("name", ["all_A", "tandem", "long_repeats", "planted"])
def test_native_builder_repetitive(name):
    rng = np.random.default_rng(3)
    T = {"all_A": np.zeros(3000, dtype=np.uint8),
         "tandem": np.tile(np.array([0, 1, 0, 1, 2], dtype=np.uint8), 777),
         "long_repeats": np.tile(rng.integers(0, 4, size=1500, dtype=np.uint8), 9),
         "planted": sim.make_genome(300000, seed=5, repeat_frac=0.2)}[name]
    same_index(ix.build_index_native(T), ix.build_index(T, device="cpu"))


@pytest.mark.skipif(not os.path.exists(R.REF_BWA), reason="oracle/_ref/bwa not built")
@pytest.mark.parametrize("algo,n", [("is", 128 * 33), ("is", 1000003), ("bwtsw", 12000017)])
def test_native_index_files_byte_identical_to_reference(tmp_path, algo, n):
    """`bwa index -a is` and `-a bwtsw` (the builder the reference needs above 2 Gb, bwtindex.c:103-106) write the same bytes."""
    T = sim.make_genome(n, seed=n)
    sim.write_fasta(str(tmp_path / "g.fa"), T, 3)
    subprocess.run([R.REF_BWA, "index", "-a", algo, "-p", str(tmp_path / "ref"), str(tmp_path / "g.fa")], check=True, capture_output=True)
    idx = ix.build_index_native(T, write_prefix=str(tmp_path / "mine"))
    for ext in ("bwt", "rbwt", "sa", "rsa"):
        assert (tmp_path / f"ref.{ext}").read_bytes() == (tmp_path / f"mine.{ext}").read_bytes(), ext
    ix.save_index(str(tmp_path / "py"), idx)  # the python writers agree with bwa_gpu_index_write
    for ext in ("bwt", "rbwt", "sa", "rsa", "pac"):
        assert (tmp_path / f"ref.{ext}").read_bytes() == (tmp_path / f"py.{ext}").read_bytes(), ext


def test_native_builder_20Mb_equals_torch_on_device():
    T = sim.make_genome(20_000_000, seed=1, repeat_frac=0.01)
    a = ix.build_index_native(T)
    os.environ["BWAGPU_TORCH_INDEX"] = "1"
    try:
        b = ix.build_index(T, device="cuda:0")
    finally:
        os.environ["BWAGPU_TORCH_INDEX"] = "0"
    same_index(a, b)


# ------------------------------------------------------------------ C4: > 3.0e9 bases
C4_BP = int(os.environ.get("BWAGPU_C4_BP", "3050000000"))


def _enough_memory():
    import torch
    free, _ = torch.cuda.mem_get_info()
    try:
        host = os.sysconf("SC_PHYS_PAGES") * os.sysconf("SC_PAGE_SIZE")
    except (ValueError, OSError):
        host = 0
    return free > 110e9 and host > 60e9


@pytest.fixture(scope="module")
def c4():
    if not _enough_memory():
        pytest.skip("C4 needs > 110 GB of free HBM and > 60 GB of host memory")
    T, idx = R.bwa.workload.genome_and_index(C4_BP, seed=1, device=0)
    api.init([0])
    api.load_index(idx)
    yield T, idx
    api.destroy()


def test_c4_index_is_the_bwt_of_the_genome(c4):
    """An index nobody else built: check it against the genome itself.  Error-free reads must be found with zero
    differences, and the SA row of a unique one must convert to the position it was cut from."""
    T, idx = c4
    rng = np.random.default_rng(7)
    n, L = 20000, 100
    pos = np.concatenate([rng.integers(0, C4_BP - L, size=n - 2000), rng.integers(C4_BP - (1 << 28), C4_BP - L, size=2000)])
    bases = T[pos[:, None] + np.arange(L)[None, :]]
    strand = rng.random(n) < 0.5
    bases = np.where(strand[:, None], 3 - bases[:, ::-1], bases).astype(np.uint8)
    offs = np.arange(n + 1, dtype=np.int64) * L
    opt = abi.default_gap_opt()
    n_aln, _, aln_off, aln = api.aln_flat(bases.reshape(-1), offs, opt)
    assert (n_aln >= 1).all()
    first = aln[aln_off[:-1]]
    assert ((first["info"] & 0xFFFFFF) == 0).all(), "an error-free read must have a zero-difference hit first"
    uniq = (n_aln == 1) & (first["k"] == first["l"])
    assert uniq.sum() > 0.9 * n
    a = ((first["info"] >> 24) & 1).astype(np.uint8)  # bwt_aln1_t.a: the strand bwa_cal_pac_pos_core switches on (bwase.c:144)
    sa = api.cal_pac_pos(first["k"][uniq], (a[uniq] != 0).astype(np.uint8))
    got = np.where(a[uniq] != 0, sa, (np.uint32(C4_BP) - (sa + np.uint32(L))).astype(np.uint32)).astype(np.int64)
    assert np.array_equal(got, pos[uniq]), "SA -> coordinate does not give back where the read came from"
    assert (pos[uniq] > (1 << 31)).sum() > 1000, "the test must exercise coordinates above 2^31"


@pytest.mark.skipif(not R.have_ref(), reason="oracle/_ref not built")
def test_c4_search_and_sa_match_reference_live(c4):
    T, idx = c4
    import torch
    ridx = R.RefIndex(idx)
    opt = abi.default_gap_opt()
    # reads cut on the host from the top of the genome and across it (simulate_reads wants the genome on the device: 3 GB, fine)
    reads = sim.simulate_reads(T, 6000, 100, seed=33, device="cuda:0")
    got = api.aln_flat(reads.bases, reads.offs, opt)
    want = R.ref_aln(ridx, reads, opt, threads=os.cpu_count() or 8)
    errs = R.compare_aln(want, got, "c4/se100")
    assert not errs, errs[:5]
    assert int((got[3]["l"] >= (1 << 31)).sum()) > 100, "SA intervals above 2^31 must occur"
    torch.cuda.empty_cache()
    rng = np.random.default_rng(5)
    k = np.concatenate([rng.integers(1, C4_BP + 1, size=60000), np.array([1, C4_BP, idx.bwt[0].primary, idx.bwt[1].primary, C4_BP - 1])]).astype(np.uint32)
    which = rng.integers(0, 2, size=k.size).astype(np.uint8)
    assert np.array_equal(api.cal_pac_pos(k, which), R.ref_sa(ridx, k, which, threads=os.cpu_count() or 8))
