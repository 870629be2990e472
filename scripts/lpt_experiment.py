"""Experiment: how much of k_search's tail would a longest-first job order remove?
Runs the resident batch once, re-stages the same reads sorted by max_entries (a proxy for search
cost) descending / ascending / random, and prints the device time of each."""
import importlib, sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
bwa = importlib.import_module("network-aware-bwa_b200")
import bench
api, abi = bwa.api, bwa.abi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
T, idx, reads = bench.make_workload(bwa, n, "cuda:0", seed=1000, genome_bp=100_000_000)
opt = abi.default_gap_opt()
api.init([0]); api.load_index(idx)
L = 76
api.resident_stage(reads.bases, reads.offs, opt)
for _ in range(2): api.resident_run()
t0 = min(api.resident_run() for _ in range(3)); st = api.get_stats()
n_aln, max_entries, off, aln = api.resident_fetch(n)
print("as simulated: %.1f ms (search %.1f)" % (t0, st["ms_search"]))
print("max_entries percentiles", np.percentile(max_entries, [50, 90, 99, 99.9, 99.99, 100]))
B = reads.bases.reshape(n, L)
for name, order in (("hard first", np.argsort(-max_entries, kind="stable")), ("easy first", np.argsort(max_entries, kind="stable")),
                    ("hardest 2% first", None)):
    if order is None:
        thr = np.percentile(max_entries, 98)
        hard = np.nonzero(max_entries >= thr)[0]; rest = np.nonzero(max_entries < thr)[0]
        order = np.concatenate([hard, rest])
    b2 = np.ascontiguousarray(B[order]).reshape(-1)
    api.resident_stage(b2, reads.offs, opt)
    api.resident_run()
    t = min(api.resident_run() for _ in range(3)); st = api.get_stats()
    print("%s: %.1f ms (search %.1f)" % (name, t, st["ms_search"]))
api.destroy()
