#!/usr/bin/env python
"""bench.py's end-to-end step (one whole bam2bam run over the shard, in-process) under several settings of the shim / library,
one process, one workload: `name:ENV=V,ENV2=V2` per variant, the device set up again between variants.

    python scripts/e2e_variants.py --steps 4 base: g3l6:BWAGPU_CALL_GROUPS=3,BWAGPU_LANES=6
"""
import argparse
import importlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("variants", nargs="+")
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--pairs", type=int, default=bench.PAIRS_PER_STEP)
    ap.add_argument("--genome-bp", type=int, default=bench.GENOME_BP)
    a = ap.parse_args()
    import torch
    bwa = importlib.import_module("network-aware-bwa_b200")
    threads = os.cpu_count() or 8
    os.environ.setdefault("BWAGPU_DEVICE", "0")
    os.environ.setdefault("BWAGPU_SHIM_THREADS", str(threads))
    os.environ.setdefault("BWAGPU_HOST_THREADS", str(max(1, min(8, threads // 3))))
    os.environ.setdefault("BWAGPU_BATCH_RECORDS", str(1 << 17))
    prefix = bwa.workload.ensure_genome_files(a.genome_bp, 1, 0)
    bam = bench.pairs_bam(bwa, prefix, lambda: bwa.workload.load_genome(prefix), a.pairs, bench.READ_LEN, 1000, "cuda:0")
    outs = bench.Outputs(f"{bam[:-4]}.out")
    torch.cuda.empty_cache()
    host = bench.Host()
    base_env = dict(os.environ)
    first = True
    for spec in a.variants:
        name, _, kv = spec.partition(":")
        env = dict(x.split("=", 1) for x in kv.split(",") if x)
        for k in list(os.environ):
            if k.startswith("BWAGPU_") and k not in base_env:
                del os.environ[k]
        os.environ.update({k: v for k, v in base_env.items() if k.startswith("BWAGPU_")})
        os.environ.update(env)
        if not first:
            host.H.bwa_gpu_batch_reset_device()
        first = False
        saved = bench.quiet_stderr(0)
        try:
            for _ in range(a.warmup):
                host.run(prefix, bam, outs.next())
            ms, reps = [], []
            for _ in range(a.steps):
                t1 = time.perf_counter()
                rep = host.run(prefix, bam, outs.next())
                ms.append(round((time.perf_counter() - t1 - rep["index_load_s"]) * 1e3, 1)); reps.append(rep)
        finally:
            bench.restore_stderr(saved)
        last = reps[-1]
        print(json.dumps({"variant": name, "env": env, "ms_each_step": ms, "reads_per_s": round(2 * a.pairs * a.steps / (sum(ms) / 1e3)),
                          "pass1_s": round(last["pass1_s"], 3), "pass2_s": round(last["pass2_s"], 3), "dev_aln_s": round(last["dev_aln_s"], 3),
                          "process_cpu_s": round(last["process_cpu_s"], 2)}), flush=True)
    host.close()
    outs.close()


if __name__ == "__main__":
    main()
