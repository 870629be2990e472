"""Generates tests/golden/aln_golden.npz by RUNNING THE REFERENCE ITSELF (oracle/_ref, built
from /root/reference by oracle/Makefile) on seeded synthetic inputs.  Run here (where the
reference exists); the .npz is committed so the GPU box needs neither /root/reference nor
oracle/_ref to pin parity.

    python tests/golden/make_golden.py

Contents: genome (uint8 0..3); per config c in CONFIGS: c_bases, c_offs (reads), c_opt (the 16
gap_opt_t words), c_n_aln, c_max_entries, c_aln (bwt_aln1_t bytes as u32[n,4]) from
bwa_cal_sa_reg_gap(bwt, 1, &seq, opt) (bwtaln.c:93), plus sa_k, sa_which, sa_out from bwt_sa
(bwt.c:72), maxdiff tables from bwa_cal_maxdiff (bwtaln.c:37), sw_*/swp_* = Smith-Waterman jobs with
(score, start_i, start_j, end_i, end_j) from aln_local_core (stdaln.c:529), swpath_* = the same call with its path
(end points + CIGAR), glob_* = aln_global_core (stdaln.c:345) jobs with (gap_end, band) per job.
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import refload as R  # noqa: E402

bwa, abi = R.bwa, R.abi

CONFIGS = {
    # name: (read length, simulate kwargs, gap_opt overrides)
    "se36": (36, {}, {}),
    "se76": (76, {}, {}),
    "pe100": (100, {}, {}),
    "adna": ((30, 50), dict(adna=True, sub_rate=0.01), dict(seed_len=1024, fnr=0.01, max_gapo=2)),
    "ragged": ((1, 140), dict(n_rate=0.02), {}),
    "fixed_n3": (50, {}, dict(fnr=-1.0, max_diff=3)),
    "nonstop_loggap": (40, {}, dict(mode=abi.BWA_MODE_GAPE | abi.BWA_MODE_COMPREAD | abi.BWA_MODE_NONSTOP | abi.BWA_MODE_LOGGAP,
                                    max_gapo=2)),
    "nogape": (60, dict(indel_frac=0.5), dict(mode=abi.BWA_MODE_COMPREAD, max_gape=3)),
}
N_READS = 1500


def opt_words(opt):
    return np.frombuffer(bytes(opt), dtype=np.uint32).copy()


def main():
    T = bwa.simulate.make_genome(300000, seed=5, repeat_frac=0.03)
    idx = bwa.index.build_index(T)
    ridx = R.RefIndex(idx)
    out = {"genome": T}
    for name, (length, kw, optkw) in CONFIGS.items():
        reads = bwa.simulate.simulate_reads(T, N_READS, length, seed=abs(hash(name)) % 1000 + 1 if False else len(name) * 7 + 3, **kw)
        if name == "ragged":  # edge cases: empty read, all-N read, homopolymers
            bases = reads.bases.copy()
            offs = reads.offs.copy()
            l0 = offs[1] - offs[0]
            bases[offs[1]:offs[2]] = 4
            bases[offs[2]:offs[3]] = 0
            bases[offs[3]:offs[4]] = 3
            reads = bwa.simulate.Reads(np.concatenate([bases, np.zeros(0, np.uint8)]),
                                       np.concatenate([offs, [offs[-1]]]).astype(np.int64),  # trailing empty read
                                       np.append(reads.pos, 0), np.append(reads.strand, False))
        opt = abi.default_gap_opt(**optkw)
        n_aln, max_entries, aln_off, aln = R.ref_aln(ridx, reads, opt)
        out[f"{name}_bases"] = reads.bases
        out[f"{name}_offs"] = reads.offs
        out[f"{name}_opt"] = opt_words(opt)
        out[f"{name}_n_aln"] = n_aln
        out[f"{name}_max_entries"] = max_entries
        out[f"{name}_aln"] = aln.view(np.uint32).reshape(-1, 4)
        print(name, "reads", reads.n, "alns", aln.size, "mean max_entries", max_entries.mean(), "max", max_entries.max())
    # bwt_sa
    rng = np.random.default_rng(3)
    n = idx.bwt[0].seq_len
    k = rng.integers(0, n + 1, size=20000, dtype=np.uint32)
    k[:40] = np.arange(40)
    k[40:80] = n - np.arange(40)
    k[80] = idx.bwt[0].primary
    k[81] = idx.bwt[1].primary
    which = rng.integers(0, 2, size=k.size, dtype=np.uint8)
    out["sa_k"], out["sa_which"], out["sa_out"] = k, which, R.ref_sa(ridx, k, which)
    L, _ = R.ref()
    out["maxdiff_004"] = np.array([L.bwa_cal_maxdiff(l, 0.02, float(np.float32(0.04))) for l in range(0, 400)], dtype=np.int32)
    out["maxdiff_001"] = np.array([L.bwa_cal_maxdiff(l, 0.02, float(np.float32(0.01))) for l in range(0, 400)], dtype=np.int32)
    # aln_local_core forward score + start/end cells (stdaln.c:529-712) on mate-rescue-like jobs
    refs, ro, qs, qo = R.make_sw_jobs(T, 2000, seed=17)
    out["sw_refs"], out["sw_ref_off"], out["sw_queries"], out["sw_q_off"] = refs, ro, qs, qo
    out["sw_out"] = R.ref_sw_batch(refs, ro, qs, qo)
    # the same on windows of the packed genome (what bwa_sw_core feeds it, bwape.c:447-456): K5's input
    refs, ro, qs, qo, begs = R.make_sw_jobs(T, 3000, seed=23, ref_n=False, with_beg=True, read_len=(25, 150), win=(30, 700))
    out["swp_beg"], out["swp_reglen"], out["swp_queries"], out["swp_q_off"] = begs, np.diff(ro).astype(np.int32), qs, qo
    out["swp_out"] = R.ref_sw_batch(refs, ro, qs, qo)
    # the full aln_local_core call (third pass included): score, path end points, CIGAR -- on the swp jobs
    rows, cigs, coff = [], [], [0]
    for i in range(begs.size):
        r = R.ref_sw_path(refs[ro[i]:ro[i + 1]], qs[qo[i]:qo[i + 1]])
        rows.append(r[:5]); cigs.append(r[5]); coff.append(coff[-1] + r[5].size)
    out["swpath_out"] = np.array(rows, dtype=np.int32)
    out["swpath_cigar"] = np.concatenate(cigs).astype(np.uint16)
    out["swpath_cigar_off"] = np.array(coff, dtype=np.int64)
    # aln_global_core as refine_gapped_core calls it (gap_end 5, band 50) and with a narrow band / free ends
    rng = np.random.default_rng(9)
    gl_beg, gl_len, gl_q, gl_qoff, gl_par, rows, cigs, coff = [], [], [], [0], [], [], [], [0]
    for i in range(1500):
        q = qs[qo[i]:qo[i + 1]][:150]
        rl = int(q.size + rng.integers(0, 9)); beg = int(begs[i] + rng.integers(0, 10))
        ge, band = [(5, 50), (-1, 50), (5, 7), (-1, 2)][i % 4]
        r = R.ref_global(T[beg:beg + rl], q, ge, band)
        gl_beg.append(beg); gl_len.append(rl); gl_q.append(q); gl_qoff.append(gl_qoff[-1] + q.size); gl_par.append((ge, band))
        rows.append(r[:5]); cigs.append(r[5]); coff.append(coff[-1] + r[5].size)
    out["glob_beg"], out["glob_reglen"] = np.array(gl_beg, dtype=np.int64), np.array(gl_len, dtype=np.int32)
    out["glob_queries"], out["glob_q_off"] = np.concatenate(gl_q).astype(np.uint8), np.array(gl_qoff, dtype=np.int64)
    out["glob_par"] = np.array(gl_par, dtype=np.int32)
    out["glob_out"] = np.array(rows, dtype=np.int32)
    out["glob_cigar"] = np.concatenate(cigs).astype(np.uint16)
    out["glob_cigar_off"] = np.array(coff, dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "aln_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "aln_golden.npz"), os.path.getsize(os.path.join(HERE, "aln_golden.npz")))


if __name__ == "__main__":
    main()
