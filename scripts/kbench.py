#!/usr/bin/env python
"""Kernel-only A/B harness: K2 + K3 on a resident batch (bwa_gpu_resident_run), device-timed; one JSON line.
    python scripts/kbench.py --genome-bp 3100000000 --read-len 100 --reads 4000000 [--batches 131072,1000000] [--check 5000]
Environment switches of the library (BWAGPU_POP_CAP, BWAGPU_LIB=variants/..., ...) apply.  TEST/BENCH INFRASTRUCTURE."""
import argparse
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
bwa = importlib.import_module("network-aware-bwa_b200")
api, abi = bwa.api, bwa.abi


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genome-bp", type=int, default=100_000_000)
    ap.add_argument("--read-len", type=int, default=76)
    ap.add_argument("--reads", type=int, default=10_000_000)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--batches", default="", help="also time these smaller resident batches (comma separated read counts)")
    ap.add_argument("--check", type=int, default=0, help="compare this many reads with the live reference")
    ap.add_argument("--stats", action="store_true")
    ap.add_argument("--tag", default="")
    a = ap.parse_args()
    if a.genome_bp >= 1_000_000_000:
        T, idx = bwa.workload.genome_and_index(a.genome_bp, 1, 0)
    else:
        T = bwa.simulate.make_genome(a.genome_bp, seed=1, repeat_frac=0.01)
        idx = bwa.index.build_index(T, device="cuda:0")
    reads = bwa.simulate.simulate_reads(T, a.reads, a.read_len, seed=1000, device="cuda:0")
    import torch
    torch.cuda.empty_cache()
    opt = abi.default_gap_opt()
    api.init([0])
    api.load_index(idx)
    out = {"tag": a.tag, "genome_bp": a.genome_bp, "read_len": a.read_len, "env": {k: v for k, v in os.environ.items() if k.startswith("BWAGPU_")}}

    def timed(n):
        sub_b, sub_o = reads.bases[: reads.offs[n]], reads.offs[: n + 1]
        api.resident_stage(sub_b, sub_o, opt)
        api.resident_run()
        ms = s = w = 0.0
        tiers = [0.0] * 4
        for _ in range(a.reps):
            ms += api.resident_run()
            st = api.get_stats()
            s += st["ms_search"]; w += st["ms_width"]
            tiers = [x + y for x, y in zip(tiers, st["ms_tier"])]
        return {"reads": n, "ms_total": ms / a.reps, "ms_search": s / a.reps, "ms_width": w / a.reps, "ms_pass": [t / a.reps for t in tiers[:3]],
                "to_pass1": int(st["n_overflow_t2"]), "to_pass2": int(st["n_overflow_t3"]), "mreads_per_s": n / (ms / a.reps) / 1e3}

    out["full"] = timed(a.reads)
    if a.stats:
        api.set_stats(True)
        api.resident_run()
        st = api.get_stats()
        api.set_stats(False)
        out["per_read"] = {k: st[k] / a.reads for k in ("occ_fetches_search", "own_fetches_search", "n_pops", "n_stored", "n_trips", "n_expand", "n_exact", "n_derive")}
        out["stats_pass_ms"] = {"queue_empty": st["ns_queue_empty"] / 1e6, "kernel": st["ns_kernel"] / 1e6}
    if a.check:
        import refload as R
        api.resident_stage(reads.bases[: reads.offs[a.check]], reads.offs[: a.check + 1], opt)
        api.resident_run()
        got = api.resident_fetch(a.check)
        sub = bwa.simulate.Reads(reads.bases[: reads.offs[a.check]], reads.offs[: a.check + 1], None, None)
        want = R.ref_aln(R.RefIndex(idx), sub, opt, threads=os.cpu_count() or 8)
        out["mismatches"] = len(R.compare_aln(want, got, "kbench"))
    out["batches"] = [timed(int(x)) for x in a.batches.split(",") if x]
    print(json.dumps(out))
    api.destroy()


if __name__ == "__main__":
    main()
