"""FM-index builder and file formats against the reference's `bwa index -a is`."""
import os
import subprocess

import numpy as np
import pytest

import refload as R

ix = R.bwa.index
sim = R.bwa.simulate


@pytest.mark.skipif(not os.path.exists(R.REF_BWA), reason="oracle/_ref/bwa not built")
@pytest.mark.parametrize("n", [257, 128 * 33, 100003])
def test_index_files_byte_identical_to_reference(tmp_path, n):
    T = sim.make_genome(n, seed=n)
    sim.write_fasta(str(tmp_path / "g.fa"), T, 3)
    subprocess.run([R.REF_BWA, "index", "-a", "is", "-p", str(tmp_path / "ref"), str(tmp_path / "g.fa")],
                   check=True, capture_output=True)
    ix.save_index(str(tmp_path / "mine"), ix.build_index(T))
    for ext in ("bwt", "rbwt", "sa", "rsa", "pac"):
        a = (tmp_path / f"ref.{ext}").read_bytes()
        b = (tmp_path / f"mine.{ext}").read_bytes()
        assert a == b, ext


def test_index_roundtrip(tmp_path):
    T = sim.make_genome(5000, seed=2)
    idx = ix.build_index(T)
    ix.save_index(str(tmp_path / "x"), idx)
    back = ix.load_index(str(tmp_path / "x"))
    for s in range(2):
        assert back.bwt[s].primary == idx.bwt[s].primary
        assert np.array_equal(back.bwt[s].bwt, idx.bwt[s].bwt)
        assert np.array_equal(back.bwt[s].sa, idx.bwt[s].sa)
        assert np.array_equal(back.bwt[s].L2, idx.bwt[s].L2)


def test_suffix_array_small_bruteforce():
    import torch
    rng = np.random.default_rng(0)
    for n in (1, 2, 17, 300):
        T = rng.integers(0, 4, size=n, dtype=np.uint8)
        sa = ix.suffix_array(torch.from_numpy(T)).numpy()
        want = sorted(range(n), key=lambda i: bytes(T[i:] + 1))
        assert list(sa) == want


def test_suffix_array_repetitive():
    import torch
    T = np.tile(np.array([0, 1, 0, 1, 2], dtype=np.uint8), 200)
    sa = ix.suffix_array(torch.from_numpy(T)).numpy()
    want = sorted(range(T.size), key=lambda i: bytes(T[i:] + 1))
    assert list(sa) == want
