"""GPU fuzz through the C-ABI: random option sets and read shapes, bwa_gpu_aln_flat against the live reference (oracle/_ref),
with the pass-0 arena varied so that reads finish in k_search, in k_search_warp and in the guaranteed pass.

    python scripts/fuzz_gpu.py <seed> <iterations> [reads per set]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import refload as R
abi, api, bwa = R.abi, R.bwa.api, R.bwa
golden = np.load(os.path.join(ROOT, "tests", "golden", "aln_golden.npz"))
T = golden["genome"]
idx = bwa.index.build_index(T)
ridx = R.RefIndex(idx)
api.init([0]); api.load_index(idx)
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 30
n = int(sys.argv[3]) if len(sys.argv) > 3 else 2000
bad = 0
for it in range(iters):
    kw = dict(s_mm=int(rng.integers(1, 6)), s_gapo=int(rng.integers(1, 14)), s_gape=int(rng.integers(1, 7)),
              max_gapo=int(rng.integers(0, 3)), max_gape=int(rng.integers(0, 8)), indel_end_skip=int(rng.integers(0, 7)),
              max_del_occ=int(rng.integers(1, 20)), seed_len=int(rng.choice([8, 16, 32, 1024])), max_seed_diff=int(rng.integers(0, 3)),
              max_top2=int(rng.choice([0, 1, 3, 30])), max_entries=int(rng.choice([200, 5000, 2000000])),
              mode=int(rng.choice([0x01, 0x00, 0x05, 0x11, 0x15, 0x04, 0x10])) | 0x02)
    if rng.random() < 0.5:
        kw["fnr"] = float(rng.choice([0.04, 0.01, 0.1, 0.001]))
    else:
        kw["fnr"], kw["max_diff"] = -1.0, int(rng.integers(0, 5))
    opt = abi.default_gap_opt(**kw)
    lo = int(rng.integers(8, 60)); hi = lo + int(rng.integers(0, 60))
    reads = bwa.simulate.simulate_reads(T, n, (lo, hi), seed=int(rng.integers(1, 1 << 30)), sub_rate=float(rng.choice([0.0, 0.02, 0.06])),
                                        n_rate=float(rng.choice([0.0, 0.01, 0.05])))
    want = R.ref_aln(ridx, reads, opt, threads=16)
    for cap, pool in (("2048", None), ("2", None), ("16", "1")):
        os.environ["BWAGPU_T1_CAP"] = cap
        if pool: os.environ["BWAGPU_POOL_MB"] = pool
        else: os.environ.pop("BWAGPU_POOL_MB", None)
        got = api.aln_flat(reads.bases, reads.offs, opt)
        st = api.get_stats()
        errs = R.compare_aln(want, got, "fuzz")
        if errs:
            bad += 1
            print(it, "MISMATCH cap", cap, errs[:1], kw, (lo, hi), flush=True)
    print(it, "done; last: warp pass", st["n_overflow_t2"], "guaranteed", st["n_overflow_t3"], "max max_entries", int(want[1].max()), flush=True)
api.destroy()
print("FUZZ", "FAILURES" if bad else "ALL OK", bad)
