"""aDNA-option probe: how much work do the reads that fail the optimistic pass represent?"""
import importlib, sys, os, time, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
bwa = importlib.import_module("network-aware-bwa_b200")
api, abi = bwa.api, bwa.abi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 300_000
T = bwa.simulate.make_genome(100_000_000, seed=1, repeat_frac=0.01)
idx = bwa.index.build_index(T, device="cuda:0")
api.init([0]); api.load_index(idx)
reads = bwa.simulate.simulate_reads(T, n, (30, 50), seed=1000, device="cuda:0", adna=True, sub_rate=0.01)
opt = abi.default_gap_opt(seed_len=1024, fnr=0.01, max_gapo=2)
api.set_stats(True)
api.resident_stage(reads.bases, reads.offs, opt)
ms = api.resident_run(); st = api.get_stats()
na, me, off, aln = api.resident_fetch(n)
print("ms", ms, "tiers", st["ms_tier"], "retried", st["n_overflow_t2"], st["n_overflow_t3"], "chunks used", st["x_chunks_used"])
print("trips total %.3g pops %.3g stored %.3g hits %d" % (st["n_trips"], st["n_pops"], st["n_stored"], aln.size))
print("max_entries pct", np.percentile(me, [50, 90, 99, 99.9, 99.99, 100]))
print("n_aln pct", np.percentile(na, [50, 90, 99, 99.9, 99.99, 100]), "reads at the 2M limit", int((me > 2000000).sum()))
api.destroy()
