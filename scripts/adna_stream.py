"""aDNA-option throughput through the chunked, multi-lane host path (bwa_gpu_aln_flat): deep searches of one
chunk overlap the next chunks."""
import importlib, sys, os, time, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
bwa = importlib.import_module("network-aware-bwa_b200")
api, abi = bwa.api, bwa.abi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
T = bwa.simulate.make_genome(100_000_000, seed=1, repeat_frac=0.01)
idx = bwa.index.build_index(T, device="cuda:0")
api.init([0]); api.load_index(idx)
reads = bwa.simulate.simulate_reads(T, n, (30, 50), seed=1000, device="cuda:0", adna=True, sub_rate=0.01)
opt = abi.default_gap_opt(seed_len=1024, fnr=0.01, max_gapo=2)
for rep in range(2):
    t = time.perf_counter(); r = api.aln_flat(reads.bases, reads.offs, opt); dt = time.perf_counter() - t
    st = api.get_stats()
    print(f"{n} aDNA reads: {dt:.2f} s = {n / dt / 1e3:.0f} K reads/s; passes ms {[round(x) for x in st['ms_tier']]} retried {st['n_overflow_t2']}/{st['n_overflow_t3']}")
api.destroy()
