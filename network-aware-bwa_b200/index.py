"""FM-index construction and (de)serialisation in the reference's own formats.

The alignment hot path consumes the index the reference's `bwa index` writes:
`<prefix>.bwt/.rbwt` (BWT with occurrence counts interleaved every 128 bases) and
`<prefix>.sa/.rsa` (suffix array sampled every 32) -- formats defined by
bwtio.c:17-38 (dump), bwtio.c:145-200 (restore), bwtmisc.c:125-152 (occ interleave),
bwt.c:48-70 (SA sampling), bwtmisc.c:168-193 (reversed, NOT complemented, strand).

Index *construction* is off the hot path (bwtindex.c/is.c/bwt_gen are
out of scope), but tests and bench.py need indexes for synthetic genomes on a box
where the reference does not exist, so this module builds them with a prefix-doubling
suffix sort written on torch tensors (runs on the B200 for the 100 Mb+ genomes, on CPU
for the small test genomes).  The BWT of a string is unique, so the output is required
to be byte-identical to `bwa index -a is` -- tests/test_index.py checks exactly that
against oracle/_ref/bwa.
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import numpy as np
import torch

OCC_INTERVAL = 128  # bwt.h:35
SA_INTV = 32  # bwtindex.c:173,185


@dataclass
class Bwt:
    """Host image of the reference's bwt_t (bwt.h:43-59), arrays in on-disk layout."""

    primary: int
    L2: np.ndarray  # uint32[5], cumulative base counts
    seq_len: int
    bwt: np.ndarray  # uint32[bwt_size], 12-word blocks: cnt[4] + 8 words of 16 bases
    sa: np.ndarray | None = None  # uint32[n_sa]; sa[0] = 0xffffffff (bwt.c:69)
    sa_intv: int = SA_INTV

    @property
    def n_sa(self) -> int:
        return (self.seq_len + self.sa_intv) // self.sa_intv


# --------------------------------------------------------------------------- suffix sort
def suffix_array(T: torch.Tensor) -> torch.Tensor:
    """Suffix array (int64[n]) of the base string T (uint8, values 0..3), sentinel smaller
    than every base (is.c:188-196 convention: the empty suffix sorts first and is NOT in
    the returned array).  Prefix doubling with full re-sorts; O(log(max LCP)) rounds."""
    n = T.numel()
    dev = T.device
    if n == 0:
        return torch.empty(0, dtype=torch.int64, device=dev)
    K = 21  # 3 bits per symbol (0 = past the end, 1..4 = A,C,G,T) -> 63 bits
    Tp = torch.zeros(n + K, dtype=torch.int64, device=dev)
    Tp[:n] = T.to(torch.int64) + 1
    key = torch.zeros(n, dtype=torch.int64, device=dev)
    for j in range(K):
        key = (key << 3) | Tp[j : j + n]
    del Tp
    h = K
    while True:
        skey, sa = torch.sort(key)
        del key
        flags = torch.ones(n, dtype=torch.int64, device=dev)
        flags[1:] = (skey[1:] != skey[:-1]).to(torch.int64)
        del skey
        r = torch.cumsum(flags, 0)  # 1-based dense ranks in sorted order
        del flags
        nranks = int(r[-1].item())
        if nranks == n:
            return sa
        rank = torch.empty(n, dtype=torch.int64, device=dev)
        rank[sa] = r
        del r, sa
        r2 = torch.zeros(n, dtype=torch.int64, device=dev)
        if h < n:
            r2[: n - h] = rank[h:]
        if (n + 1) * (n + 1) >= 2**63:
            raise NotImplementedError("suffix_array: n too large for single-key doubling")
        key = rank * (n + 1) + r2
        del rank, r2
        h *= 2


# --------------------------------------------------------------------------- BWT + occ + SA
def _pack16(bases: torch.Tensor) -> torch.Tensor:
    """2-bit pack, 16 bases per uint32, first base in the top bits (bwtmisc.c:97-98)."""
    n = bases.numel()
    nw = (n + 15) >> 4
    buf = torch.zeros(nw * 16, dtype=torch.int64, device=bases.device)
    buf[:n] = bases.to(torch.int64)
    buf = buf.view(nw, 16)
    shifts = torch.arange(15, -1, -1, dtype=torch.int64, device=bases.device) * 2
    return (buf << shifts).sum(1)  # int64 holding a u32


def build_bwt(T: torch.Tensor) -> Bwt:
    """bwt_pac2bwt (bwtmisc.c:56-101) + bwt_bwtupdate_core (125-152) + bwt_cal_sa(32)
    (bwt.c:48-70), from a full suffix sort instead of IS + LF walking."""
    n = T.numel()
    dev = T.device
    sa = suffix_array(T)
    sa_full = torch.cat([torch.tensor([n], dtype=torch.int64, device=dev), sa])
    del sa
    primary = int(torch.nonzero(sa_full == 0)[0].item())  # is.c:209-211
    # sampled SA: sa[k/32] = SA_full[k]; sa[0] := -1 (bwt.c:62-69)
    sa_s = sa_full[::SA_INTV].clone()
    sa_s[0] = 0xFFFFFFFF
    # BWT with the '$' row removed (is.c:212-213)
    prev = sa_full - 1
    prev[primary] = 0
    B = T[prev]
    B = torch.cat([B[:primary], B[primary + 1 :]])
    del prev, sa_full
    counts = torch.bincount(T.to(torch.int64), minlength=4)[:4]
    L2 = np.zeros(5, dtype=np.uint32)
    L2[1:] = np.cumsum(counts.cpu().numpy()).astype(np.uint32)
    # occ interleave
    words = _pack16(B)  # ceil(n/16)
    n_occ = (n + OCC_INTERVAL - 1) // OCC_INTERVAL + 1
    onehot = torch.zeros((n_occ - 1) * OCC_INTERVAL, dtype=torch.int64, device=dev)
    cum = torch.zeros((n_occ, 4), dtype=torch.int64, device=dev)
    for c in range(4):
        onehot.zero_()
        onehot[:n] = (B == c).to(torch.int64)
        cum[1:, c] = torch.cumsum(onehot.view(n_occ - 1, OCC_INTERVAL).sum(1), 0)
    del onehot
    nw = words.numel()
    bwt_size = nw + 4 * n_occ
    out = torch.zeros(bwt_size, dtype=torch.int64, device=dev)
    widx = torch.arange(nw, dtype=torch.int64, device=dev)
    out[widx + 4 * (widx // 8 + 1)] = words  # word w sits after (w/8 + 1) count groups
    full = n // OCC_INTERVAL  # count groups 0..full sit on 12-word boundaries
    nb = full + 1 if n % OCC_INTERVAL else full
    # groups whose 128-base block starts inside the string: b = 0 .. ceil(n/128)-1
    b = torch.arange(nb, dtype=torch.int64, device=dev)
    base = b * 12
    for c in range(4):
        out[base + c] = cum[:nb, c]
    # the trailing group (total counts) follows the last emitted word (bwtmisc.c:140-141)
    out[bwt_size - 4 : bwt_size] = cum[n_occ - 1]
    if n % OCC_INTERVAL == 0 and nb < n_occ:
        pass  # trailing group already lands on the 12-word boundary == bwt_size-4
    return Bwt(
        primary=primary,
        L2=L2,
        seq_len=n,
        bwt=out.cpu().numpy().astype(np.uint32),
        sa=sa_s.cpu().numpy().astype(np.uint32),
    )


@dataclass
class FMIndex:
    """Both strands' indexes + the packed forward sequence, as bam2bam holds them
    (globals bwt[2], pac: bam2bam.c:88-92; bwt[0] = forward .bwt/.sa, bwt[1] = .rbwt/.rsa)."""

    bwt: list  # [Bwt forward, Bwt reverse]
    pac: np.ndarray  # uint8, 4 bases per byte, first base in the top bits (bntseq.c)
    l_pac: int


def pack_pac(T: np.ndarray) -> np.ndarray:
    n = T.size
    nb = n // 4 + 1  # bwt_restore_pac reads l_pac/4+1 bytes (bwtio.c:149)
    out = np.zeros(nb, dtype=np.uint8)
    full = n // 4
    q = T[: full * 4].reshape(full, 4)
    np.left_shift(q[:, 0], 6, out=out[:full])
    out[:full] |= q[:, 1] << 4
    out[:full] |= q[:, 2] << 2
    out[:full] |= q[:, 3]
    for j in range(n - full * 4):  # the ragged tail
        out[full] |= np.uint8(int(T[full * 4 + j]) << (6 - 2 * j))
    return out


def build_index_native(T: np.ndarray, device: int = 0, write_prefix: str | None = None) -> FMIndex:
    """The library's own builder (csrc/indexbuild.cu, bwa_gpu_index_build): suffix sort on the device with cub radix
    sorts, any genome the 32-bit bwtint_t admits.  Needs a CUDA device."""
    from . import api
    pac = pack_pac(T)
    fwd, rev = api.index_build(pac, int(T.size), device=device, write_prefix=write_prefix)
    return FMIndex(bwt=[fwd, rev], pac=pac, l_pac=int(T.size))


def build_index(T: np.ndarray, device: str | torch.device = "cpu") -> FMIndex:
    """Index a base string (uint8 0..3; callers resolve N beforehand like bntseq.c:225).  On a CUDA device this is the
    library's native builder; the torch prefix-doubling sort below is the CPU harness path of the test suite."""
    dev = torch.device(device)
    if dev.type == "cuda" and os.environ.get("BWAGPU_TORCH_INDEX", "0") == "0":
        return build_index_native(T, device=dev.index or 0)
    t = torch.from_numpy(np.ascontiguousarray(T)).to(device)
    fwd = build_bwt(t)
    rev = build_bwt(torch.flip(t, [0]))
    return FMIndex(bwt=[fwd, rev], pac=pack_pac(T), l_pac=int(T.size))


# --------------------------------------------------------------------------- file formats
def dump_bwt(fn: str, b: Bwt) -> None:
    """bwt_dump_bwt (bwtio.c:17-25)."""
    with open(fn, "wb") as f:
        np.array([b.primary], dtype=np.uint32).tofile(f)
        b.L2[1:].astype(np.uint32).tofile(f)
        b.bwt.astype(np.uint32).tofile(f)


def dump_sa(fn: str, b: Bwt) -> None:
    """bwt_dump_sa (bwtio.c:27-38): sa[0] is not stored."""
    with open(fn, "wb") as f:
        np.array([b.primary], dtype=np.uint32).tofile(f)
        b.L2[1:].astype(np.uint32).tofile(f)
        np.array([b.sa_intv, b.seq_len], dtype=np.uint32).tofile(f)
        b.sa[1:].astype(np.uint32).tofile(f)


def restore_bwt(fn: str) -> Bwt:
    """bwt_restore_bwt (bwtio.c:180-200)."""
    raw = np.fromfile(fn, dtype=np.uint32)
    L2 = np.zeros(5, dtype=np.uint32)
    L2[1:] = raw[1:5]
    return Bwt(primary=int(raw[0]), L2=L2, seq_len=int(L2[4]), bwt=raw[5:].copy())


def restore_sa(fn: str, b: Bwt) -> None:
    """bwt_restore_sa (bwtio.c:157-178)."""
    raw = np.fromfile(fn, dtype=np.uint32)
    if int(raw[0]) != b.primary:
        raise ValueError("SA-BWT inconsistency: primary is not the same.")
    b.sa_intv = int(raw[5])
    if int(raw[6]) != b.seq_len:
        raise ValueError("SA-BWT inconsistency: seq_len is not the same.")
    sa = np.empty(b.n_sa, dtype=np.uint32)
    sa[0] = 0xFFFFFFFF
    sa[1:] = raw[7 : 7 + b.n_sa - 1]
    b.sa = sa


def save_index(prefix: str, idx: FMIndex) -> None:
    dump_bwt(prefix + ".bwt", idx.bwt[0])
    dump_bwt(prefix + ".rbwt", idx.bwt[1])
    dump_sa(prefix + ".sa", idx.bwt[0])
    dump_sa(prefix + ".rsa", idx.bwt[1])
    with open(prefix + ".pac", "wb") as f:  # bntseq.c:236-246: packed bytes, [0 if len%4==0], len%4
        n = idx.l_pac
        body = idx.pac[: (n >> 2) + (1 if n & 3 else 0)]
        body.tofile(f)
        if n % 4 == 0:
            f.write(b"\0")
        f.write(bytes([n % 4]))


def load_index(prefix: str, with_pac: bool = True) -> FMIndex:
    fwd = restore_bwt(prefix + ".bwt")
    rev = restore_bwt(prefix + ".rbwt")
    restore_sa(prefix + ".sa", fwd)
    restore_sa(prefix + ".rsa", rev)
    pac = np.zeros(1, dtype=np.uint8)
    if with_pac and os.path.exists(prefix + ".pac"):
        pac = np.fromfile(prefix + ".pac", dtype=np.uint8)[: fwd.seq_len // 4 + 1]
    return FMIndex(bwt=[fwd, rev], pac=pac, l_pac=fwd.seq_len)
