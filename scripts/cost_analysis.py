"""Which reads are expensive?  max_entries (search-cost proxy) by Hamming distance of the read to its
simulated origin (>= 6 ~ indel or junk)."""
import importlib, sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
bwa = importlib.import_module("network-aware-bwa_b200")
import bench
api, abi = bwa.api, bwa.abi
n = 1_000_000
T, idx, reads = bench.make_workload(bwa, n, "cuda:0", seed=1000, genome_bp=100_000_000)
opt = abi.default_gap_opt()
api.init([0]); api.load_index(idx)
L = 76
api.set_stats(False)
n_aln, max_entries, off, aln = api.aln_flat(reads.bases, reads.offs, opt)
B = reads.bases.reshape(n, L)
orig = T[reads.pos[:, None] + np.arange(L)[None, :]]
rc = (3 - B[:, ::-1])
fwd = np.where(reads.strand[:, None], rc, B)
ham = (fwd != orig).sum(1)
first_score = np.full(n, -1); has = n_aln > 0
first_score[has] = aln["score"][off[:-1][has]]
for h in range(0, 8):
    sel = ham == h if h < 7 else ham >= 7
    if sel.sum() == 0: continue
    print(f"hamming {'>=7' if h == 7 else h}: n={sel.sum():7d}  mean max_entries={max_entries[sel].mean():8.1f}  p90={np.percentile(max_entries[sel], 90):7.0f}  mapped={(n_aln[sel] > 0).mean():.3f}")
for sc in sorted(set(first_score.tolist())):
    sel = first_score == sc
    print(f"best score {sc:3d}: n={sel.sum():7d} mean max_entries={max_entries[sel].mean():8.1f}")
api.destroy()
