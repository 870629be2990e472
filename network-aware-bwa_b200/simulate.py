"""Seeded synthetic genomes and simulated reads of the shapes BASELINE.json names.

SURVEY.md §8(d): i.i.d. uniform ACGT contigs with planted repeat families (so that the
c1 > 1 / XA / repeat paths fire), reads with substitutions, short indels, a few N, and an
ancient-DNA mode (30-50 bp, C->T at the 5' end / G->A at the 3' end with p = 0.3*0.5^d).
Everything is a pure function of the seed, and is written on torch tensors so the 10 M-read
bench workload is generated on the device in a second or two (CPU works too).

Base codes are the reference's: 0..3 = A,C,G,T, 4 = N (bwaseqio.c:10; bntseq.c nst_nt4_table).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch


def make_genome(n: int, seed: int = 1, repeat_frac: float = 0.01, repeat_len: int = 600,
                max_copies: int = 6, max_div: float = 0.03) -> np.ndarray:
    """i.i.d. uniform genome of n bases (uint8 0..3) with ~repeat_frac of it overwritten by
    copies of repeat families (2..max_copies copies, 0..max_div substitution divergence)."""
    rng = np.random.default_rng(seed)
    T = rng.integers(0, 4, size=n, dtype=np.uint8)
    if repeat_frac > 0 and n > 4 * repeat_len:
        budget = int(n * repeat_frac)
        while budget > 0:
            copies = int(rng.integers(2, max_copies + 1))
            L = int(min(repeat_len, n // 8))
            src = int(rng.integers(0, n - L))
            fam = T[src : src + L].copy()
            for _ in range(copies - 1):
                dst = int(rng.integers(0, n - L))
                cp = fam.copy()
                div = rng.uniform(0, max_div)
                m = rng.random(L) < div
                cp[m] = (cp[m] + rng.integers(1, 4, size=int(m.sum()), dtype=np.uint8)) & 3
                T[dst : dst + L] = cp
                budget -= L
    return T


def write_fasta(path: str, T: np.ndarray, n_contigs: int = 4) -> list:
    """Write T as n_contigs FASTA records (60 columns); returns [(name, length)]."""
    n = T.size
    bounds = [n * i // n_contigs for i in range(n_contigs + 1)]
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    out = []
    with open(path, "wb") as f:
        for c in range(n_contigs):
            seg = lut[T[bounds[c] : bounds[c + 1]]]
            name = f"chr{c + 1}"
            out.append((name, int(seg.size)))
            f.write(f">{name}\n".encode())
            full = seg.size // 60 * 60
            if full:
                body = np.empty((full // 60, 61), dtype=np.uint8)
                body[:, :60] = seg[:full].reshape(-1, 60)
                body[:, 60] = 10
                f.write(body.tobytes())
            if seg.size > full:
                f.write(seg[full:].tobytes() + b"\n")
    return out


@dataclass
class Reads:
    """Flat read batch: bases[offs[i]:offs[i+1]] is read i in sequencing orientation
    (uint8, 0..4).  `pos`/`strand` are the simulated truth (not used by the aligner)."""

    bases: np.ndarray
    offs: np.ndarray  # int64[n+1]
    pos: np.ndarray
    strand: np.ndarray

    @property
    def n(self) -> int:
        return self.offs.size - 1

    def lens(self) -> np.ndarray:
        return np.diff(self.offs).astype(np.int32)

    def read(self, i: int) -> np.ndarray:
        return self.bases[self.offs[i] : self.offs[i + 1]]


def simulate_reads(T, n_reads: int, length, seed: int = 7, sub_rate: float = 0.015,
                   indel_frac: float = 0.1, n_rate: float = 0.001, adna: bool = False,
                   junk_frac: float = 0.01, device="cpu") -> Reads:
    """Single-end reads.  `length` is an int or an inclusive (lo, hi) range.
    indel_frac = fraction of reads carrying one 1-3 base insertion or deletion."""
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    Tt = torch.as_tensor(T).to(dev)
    n = Tt.numel()
    if isinstance(length, int):
        lo = hi = length
    else:
        lo, hi = length
    L = hi
    lens = torch.randint(lo, hi + 1, (n_reads,), generator=g, device=dev)
    start = torch.randint(0, n - L - 8, (n_reads,), generator=g, device=dev)
    idx = torch.arange(L, device=dev).unsqueeze(0).expand(n_reads, L)
    # one indel per selected read: deletion shifts the template index, insertion stalls it
    has_indel = torch.rand(n_reads, generator=g, device=dev) < indel_frac
    is_del = torch.rand(n_reads, generator=g, device=dev) < 0.5
    ilen = torch.randint(1, 4, (n_reads,), generator=g, device=dev)
    ipos = (torch.rand(n_reads, generator=g, device=dev) * (lens - 12).clamp(min=1)).long() + 6
    ipos = ipos.unsqueeze(1)
    ilen2 = ilen.unsqueeze(1)
    shift = torch.zeros((n_reads, L), dtype=torch.int64, device=dev)
    dmask = (has_indel & is_del).unsqueeze(1)
    imask = (has_indel & ~is_del).unsqueeze(1)
    shift = torch.where(dmask & (idx >= ipos), ilen2.expand(n_reads, L), shift)
    ins_here = imask & (idx >= ipos) & (idx < ipos + ilen2)
    shift = torch.where(imask & (idx >= ipos + ilen2), -ilen2.expand(n_reads, L), shift)
    tpl = start.unsqueeze(1) + idx + shift
    reads = Tt[tpl.clamp(0, n - 1)].to(torch.uint8)
    rnd_base = torch.randint(0, 4, (n_reads, L), generator=g, device=dev, dtype=torch.uint8)
    reads = torch.where(ins_here, rnd_base, reads)
    # substitutions
    sub = torch.rand((n_reads, L), generator=g, device=dev) < sub_rate
    add = torch.randint(1, 4, (n_reads, L), generator=g, device=dev, dtype=torch.uint8)
    reads = torch.where(sub, (reads + add) & 3, reads)
    # junk reads (unalignable)
    junk = torch.rand(n_reads, generator=g, device=dev) < junk_frac
    reads = torch.where(junk.unsqueeze(1), rnd_base, reads)
    # strand: reverse-complement half of them within their own length
    strand = torch.rand(n_reads, generator=g, device=dev) < 0.5
    ridx = (lens.unsqueeze(1) - 1 - idx).clamp(min=0)
    rc = 3 - torch.gather(reads, 1, ridx)
    reads = torch.where(strand.unsqueeze(1), rc, reads)
    if adna:  # deamination: C->T near the 5' end, G->A near the 3' end of the sequenced strand
        d5 = idx.float()
        d3 = (lens.unsqueeze(1) - 1 - idx).clamp(min=0).float()
        u = torch.rand((n_reads, L), generator=g, device=dev)
        ct = (reads == 1) & (u < 0.3 * torch.pow(0.5, d5))
        ga = (reads == 2) & (u < 0.3 * torch.pow(0.5, d3))
        reads = torch.where(ct, torch.full_like(reads, 3), reads)
        reads = torch.where(ga, torch.full_like(reads, 0), reads)
    # N
    nm = torch.rand((n_reads, L), generator=g, device=dev) < n_rate
    reads = torch.where(nm, torch.full_like(reads, 4), reads)
    valid = idx < lens.unsqueeze(1)
    flat = reads[valid].cpu().numpy()
    offs = np.zeros(n_reads + 1, dtype=np.int64)
    offs[1:] = np.cumsum(lens.cpu().numpy())
    return Reads(bases=flat, offs=offs, pos=start.cpu().numpy(), strand=strand.cpu().numpy())


def simulate_pairs(T, n_pairs: int, length: int, seed: int = 11, isize_mean: float = 300.0,
                   isize_sd: float = 30.0, sub_rate: float = 0.015, bad_mate_frac: float = 0.08,
                   bad_mate_sub: float = 0.10, chimeric_frac: float = 0.03, device="cpu"):
    """Paired-end FR reads: returns (Reads mate1, Reads mate2), mates at the same index.
    bad_mate_frac of the pairs have mate 2 mutated at bad_mate_sub (mate-rescue candidates);
    chimeric_frac have mate 2 drawn from an unrelated locus (discordant)."""
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    Tt = torch.as_tensor(T).to(dev)
    n = Tt.numel()
    L = length
    isz = (torch.randn(n_pairs, generator=g, device=dev) * isize_sd + isize_mean).round().long().clamp(min=L + 10)
    fdt = torch.float64 if n > (1 << 24) else torch.float32  # float32 has 24 bits: too coarse a grid for a real genome
    start = (torch.rand(n_pairs, generator=g, device=dev, dtype=fdt) * (n - isz.max().item() - 8)).long()
    idx = torch.arange(L, device=dev).unsqueeze(0)
    m1 = Tt[start.unsqueeze(1) + idx].to(torch.uint8)
    s2 = start + isz - L
    chim = torch.rand(n_pairs, generator=g, device=dev) < chimeric_frac
    s2 = torch.where(chim, (torch.rand(n_pairs, generator=g, device=dev, dtype=fdt) * (n - L - 8)).long(), s2)
    m2f = Tt[s2.unsqueeze(1) + idx].to(torch.uint8)
    m2 = 3 - torch.flip(m2f, [1])
    bad = torch.rand(n_pairs, generator=g, device=dev) < bad_mate_frac

    def mutate(r, rate):
        sub = torch.rand(r.shape, generator=g, device=dev) < rate
        add = torch.randint(1, 4, r.shape, generator=g, device=dev, dtype=torch.uint8)
        return torch.where(sub, (r + add) & 3, r)

    m1 = mutate(m1, sub_rate)
    rate2 = torch.where(bad, torch.tensor(bad_mate_sub, device=dev), torch.tensor(sub_rate, device=dev)).unsqueeze(1)
    m2 = mutate(m2, rate2.expand(n_pairs, L))
    # flip the whole fragment's strand for half the pairs (swap roles)
    flip = torch.rand(n_pairs, generator=g, device=dev) < 0.5
    a = torch.where(flip.unsqueeze(1), m2, m1)
    b = torch.where(flip.unsqueeze(1), m1, m2)
    offs = np.arange(n_pairs + 1, dtype=np.int64) * L
    r1 = Reads(a.reshape(-1).cpu().numpy(), offs, start.cpu().numpy(), flip.cpu().numpy())
    r2 = Reads(b.reshape(-1).cpu().numpy(), offs.copy(), s2.cpu().numpy(), (~flip).cpu().numpy())
    return r1, r2
