"""GPU check + timing of the warp-per-read deep pass (csrc/search_warp.cuh).

  part 1  golden vectors with a 2-record pass-0 arena: (nearly) every read goes through k_search_warp; bit-exact or it prints the diff
  part 2  live reference on option sets that stress it: aDNA options, many buckets, equal penalties (shared target slot),
          a zero penalty (serial rounds), long reads
  part 3  C5-shaped timing (aDNA reads, -l 1024 -n 0.01 -o 2, 100 Mb genome): warp pass on / off
"""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import refload as R
from test_kernel_logic import CONFIGS, golden_case
abi, api = R.abi, R.bwa.api
bwa = R.bwa
what = sys.argv[1] if len(sys.argv) > 1 else "123"
n_time = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
bad = 0

if "1" in what or "2" in what:
    golden = np.load(os.path.join(ROOT, "tests", "golden", "aln_golden.npz"))
    T = golden["genome"]
    idx = bwa.index.build_index(T)
    api.init([0]); api.load_index(idx)
    if "1" in what:
        os.environ["BWAGPU_T1_CAP"] = "2"
        for name in CONFIGS:
            reads, opt, want = golden_case(golden, name)
            got = api.aln_flat(reads.bases, reads.offs, opt)
            st = api.get_stats()
            errs = R.compare_aln(want, got, name)
            bad += len(errs)
            print(f"golden {name}: reads {reads.n} via warp pass {st['n_overflow_t2']} then guaranteed {st['n_overflow_t3']} chunks {st['x_chunks_used']}: {'OK' if not errs else errs[:3]}", flush=True)
        os.environ["BWAGPU_T1_CAP"] = "48"; os.environ["BWAGPU_POOL_MB"] = "1"; os.environ["BWAGPU_CHUNK"] = "97"
        reads, opt, want = golden_case(golden, "pe100")
        got = api.aln_flat(reads.bases, reads.offs, opt); st = api.get_stats()
        errs = R.compare_aln(want, got, "tiny pool"); bad += len(errs)
        print(f"tiny pool: warp {st['n_overflow_t2']} guaranteed {st['n_overflow_t3']}: {'OK' if not errs else errs[:3]}", flush=True)
        for k in ("BWAGPU_POOL_MB", "BWAGPU_CHUNK"): del os.environ[k]
    if "2" in what and R.have_ref():
        ridx = R.RefIndex(idx)
        cases = [
            ("adna", (30, 50), dict(seed_len=1024, fnr=0.01, max_gapo=2), dict(adna=True, sub_rate=0.01), 20000),
            ("adna deep", (30, 50), dict(seed_len=1024, fnr=0.01, max_gapo=2), dict(adna=True, sub_rate=0.06), 20000),
            ("many buckets", 60, dict(s_mm=5, s_gapo=20, s_gape=8, fnr=-1.0, max_diff=6, max_gapo=2), {}, 3000),
            ("equal penalties", 50, dict(s_mm=4, s_gapo=9, s_gape=4, max_gapo=2), dict(sub_rate=0.04), 5000),
            ("all equal", 50, dict(s_mm=3, s_gapo=3, s_gape=3, max_gapo=2), dict(sub_rate=0.04), 5000),
            ("zero gape", 50, dict(s_mm=3, s_gapo=11, s_gape=0, max_gapo=1), dict(sub_rate=0.03), 3000),
            ("ragged", (15, 250), dict(max_gapo=2, max_gape=10), dict(n_rate=0.005), 20000),
            ("low max_entries", (30, 50), dict(seed_len=1024, fnr=0.01, max_gapo=2, max_entries=3000), dict(adna=True, sub_rate=0.05), 10000),
        ]
        for cap in ("2", "2048"):
            os.environ["BWAGPU_T1_CAP"] = cap
            for name, length, optkw, simkw, n in cases:
                reads = bwa.simulate.simulate_reads(T, n, length, seed=4242, **simkw)
                opt = abi.default_gap_opt(**optkw)
                want = R.ref_aln(ridx, reads, opt, threads=16)
                t0 = time.perf_counter(); got = api.aln_flat(reads.bases, reads.offs, opt); dt = time.perf_counter() - t0
                st = api.get_stats()
                errs = R.compare_aln(want, got, name); bad += len(errs)
                print(f"live cap {cap} {name}: warp {st['n_overflow_t2']} guaranteed {st['n_overflow_t3']} passes ms {[round(x, 1) for x in st['ms_tier']]} max max_entries {int(want[1].max())}: {'OK' if not errs else errs[:3]}", flush=True)
        del os.environ["BWAGPU_T1_CAP"]
    api.destroy()

if "3" in what:
    T = bwa.simulate.make_genome(100_000_000, seed=1, repeat_frac=0.01)
    idx = bwa.index.build_index(T, device="cuda:0")
    api.init([0]); api.load_index(idx)
    reads = bwa.simulate.simulate_reads(T, n_time, (30, 50), seed=1000, device="cuda:0", adna=True, sub_rate=0.01)
    opt = abi.default_gap_opt(seed_len=1024, fnr=0.01, max_gapo=2)
    res = {}
    for warp in (("1",) if os.environ.get("BWAGPU_WARP_ONLY") else ("1", "0")):  # BWAGPU_WARP_TEAM (unset: by pass size) picks the form of the warp pass
        os.environ["BWAGPU_WARP_PASS"] = warp
        api.resident_stage(reads.bases, reads.offs, opt)
        for rep in range(2):
            ms = api.resident_run(); st = api.get_stats()
            print(f"C5 timing warp={warp}: {n_time} reads {ms:.1f} ms = {n_time / ms:.0f} K reads/s; passes ms {[round(x, 1) for x in st['ms_tier']]} retried {st['n_overflow_t2']}/{st['n_overflow_t3']} chunks {st['x_chunks_used']}", flush=True)
        res[warp] = api.resident_fetch(n_time)
    same = all(np.array_equal(a, b) for a, b in zip(res["1"], res.get("0", res["1"])))
    print("C5 results identical between warp pass and thread pass:", same, flush=True)
    bad += 0 if same else 1
    me = res["1"][1]
    print("max_entries pct", np.percentile(me, [50, 90, 99, 99.9, 99.99, 100]))
    api.destroy()
print("FAILURES" if bad else "ALL OK", bad)
