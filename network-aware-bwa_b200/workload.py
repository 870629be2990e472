"""Synthetic workloads of the shapes BASELINE.json names, as files the reference's own `bam2bam` can run on.

A workload is a seeded genome (simulate.make_genome), its index in the reference's on-disk formats (`<prefix>.bwt/.rbwt/
.sa/.rsa/.pac/.ann/.amb`, built by the library's device builder: byte-identical to `bwa index`, tests/test_index_gpu.py)
and seeded reads.  Building the 3.1 Gb index of configs[3] takes the better part of a minute even on a B200 and every arm
of bench.py (this repo's and the reference's, at every GPU count) needs the same files, so they are cached under
$BWAGPU_CACHE (default $TMPDIR/bwagpu_cache) behind a file lock: the first process on a box builds, the others wait.
Nothing here is on the timed path."""
from __future__ import annotations

import fcntl
import os
import sys
import time

import numpy as np

from . import index as ix
from . import simulate


def cache_root() -> str:
    d = os.environ.get("BWAGPU_CACHE") or os.path.join(os.environ.get("TMPDIR", "/tmp"), "bwagpu_cache")
    os.makedirs(d, exist_ok=True)
    return d


def _log(*a):
    print("[workload]", *a, file=sys.stderr, flush=True)


def write_bns(prefix: str, n: int, n_contigs: int = 4) -> None:
    """.ann / .amb as bns_dump writes them (bntseq.c:63-85) for N-free contigs cut like simulate.write_fasta."""
    bounds = [n * i // n_contigs for i in range(n_contigs + 1)]
    with open(prefix + ".ann", "w") as f:
        f.write(f"{n} {n_contigs} 11\n")
        for c in range(n_contigs):
            f.write(f"0 chr{c + 1} (null)\n{bounds[c]} {bounds[c + 1] - bounds[c]} 0\n")
    with open(prefix + ".amb", "w") as f:
        f.write(f"{n} {n_contigs} 0\n")


def write_pac(prefix: str, pac: np.ndarray, n: int) -> None:
    """bntseq.c:236-246: the packed bytes, one zero byte more when n % 4 == 0, then n % 4."""
    with open(prefix + ".pac", "wb") as f:
        pac[: (n >> 2) + (1 if n & 3 else 0)].tofile(f)
        if n % 4 == 0:
            f.write(b"\0")
        f.write(bytes([n % 4]))


def unpack_pac(pac: np.ndarray, n: int) -> np.ndarray:
    out = np.empty(((n + 3) // 4) * 4, dtype=np.uint8)
    q = out.reshape(-1, 4)
    src = pac[: q.shape[0]]
    for j in range(4):
        np.right_shift(src, 6 - 2 * j, out=q[:, j])
    out &= 3
    return out[:n]


class _Lock:
    def __init__(self, path):
        self.path = path

    def __enter__(self):
        self.f = open(self.path, "w")
        fcntl.flock(self.f, fcntl.LOCK_EX)
        return self

    def __exit__(self, *a):
        fcntl.flock(self.f, fcntl.LOCK_UN)
        self.f.close()


def genome_key(bp: int, seed: int) -> str:
    return f"g{bp}_s{seed}"


def ensure_genome_files(bp: int, seed: int = 1, device: int = 0, repeat_frac: float = 0.01) -> str:
    """-> index prefix of the cached workload genome; builds genome + index files on `device` when they are not there yet."""
    d = os.path.join(cache_root(), genome_key(bp, seed))
    os.makedirs(d, exist_ok=True)
    prefix = os.path.join(d, "g")
    done = os.path.join(d, ".done")
    with _Lock(os.path.join(d, ".lock")):
        if os.path.exists(done):
            return prefix
        t0 = time.time()
        T = simulate.make_genome(bp, seed=seed, repeat_frac=repeat_frac)
        t1 = time.time()
        pac = ix.pack_pac(T)
        import torch
        if torch.cuda.is_available():
            del T
            from . import api
            fwd, rev = api.index_build(pac, bp, device=device, write_prefix=prefix)
            del fwd, rev
        else:  # a box without a device (the CPU test suite, small genomes): the torch harness builder writes the same files
            idx = ix.build_index(T, device="cpu")
            ix.dump_bwt(prefix + ".bwt", idx.bwt[0]); ix.dump_bwt(prefix + ".rbwt", idx.bwt[1])
            ix.dump_sa(prefix + ".sa", idx.bwt[0]); ix.dump_sa(prefix + ".rsa", idx.bwt[1])
            del T, idx
        t2 = time.time()
        write_pac(prefix, pac, bp)
        write_bns(prefix, bp, 4)
        open(done, "w").write("ok\n")
        _log(f"{bp} bp genome {t1 - t0:.1f}s, index + files {time.time() - t1:.1f}s (device build {t2 - t1:.1f}s) -> {prefix}")
    return prefix


def load_genome(prefix: str) -> np.ndarray:
    n = int(open(prefix + ".ann").readline().split()[0])
    pac = np.fromfile(prefix + ".pac", dtype=np.uint8)
    return unpack_pac(pac, n)


def genome_and_index(bp: int, seed: int = 1, device: int = 0):
    """(T, FMIndex) of the cached workload genome, host arrays in the reference's layout."""
    prefix = ensure_genome_files(bp, seed, device)
    return load_genome(prefix), ix.load_index(prefix)
