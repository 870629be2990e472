"""Throughput of the other BASELINE.json read shapes (resident batch, device-timed) next to the reference's
CPU rate on a sample: C1 36 bp / 5 Mb, C3-like 100 bp / 100 Mb, C5 aDNA 30-50 bp with -l 1024 -n 0.01 -o 2."""
import importlib, sys, os, time, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
bwa = importlib.import_module("network-aware-bwa_b200")
import refload as R
api, abi = bwa.api, bwa.abi
shapes = [("C1 SE 36bp / 5 Mb", 5_000_000, 36, {}, {}, 4_000_000),
          ("C3-like 100bp / 100 Mb", 100_000_000, 100, {}, {}, 4_000_000),
          ("C5 aDNA 30-50bp / 100 Mb, -l 1024 -n 0.01 -o 2", 100_000_000, (30, 50), dict(adna=True, sub_rate=0.01), dict(seed_len=1024, fnr=0.01, max_gapo=2), 1_000_000)]
api.init([0])
cache = {}
for name, gbp, length, simkw, optkw, n in shapes:
    if gbp not in cache:
        T = bwa.simulate.make_genome(gbp, seed=1, repeat_frac=0.01)
        cache[gbp] = (T, bwa.index.build_index(T, device="cuda:0"))
    T, idx = cache[gbp]
    api.load_index(idx)
    reads = bwa.simulate.simulate_reads(T, n, length, seed=1000, device="cuda:0", **simkw)
    opt = abi.default_gap_opt(**optkw)
    api.resident_stage(reads.bases, reads.offs, opt)
    api.resident_run()
    ms = min(api.resident_run() for _ in range(2)); st = api.get_stats()
    got = api.resident_fetch(n)
    m = 30000
    sub = bwa.simulate.Reads(reads.bases[:reads.offs[m]], reads.offs[:m + 1], None, None)
    t = time.perf_counter(); want = R.ref_aln(R.RefIndex(idx), sub, opt, threads=os.cpu_count()); dt = time.perf_counter() - t
    errs = R.compare_aln(want, (got[0][:m], got[1][:m], got[2][:m + 1], got[3][:got[2][m]]), name)
    print(f"{name}: {n / ms / 1e3:.2f} M reads/s (tiers ms {[round(x, 1) for x in st['ms_tier']]}, tier2 {st['n_overflow_t2']}, tier3 {st['n_overflow_t3']}); "
          f"reference {m / dt / 1e3:.0f} K reads/s on {os.cpu_count()} cores (incl. ctypes marshalling); parity mismatches {len(errs)}")
api.destroy()
