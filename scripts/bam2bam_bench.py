#!/usr/bin/env python
"""End-to-end `bam2bam` (BAM in -> BAM out, both passes) on one box: the unmodified reference on the host
CPUs (`-t 1` sequential = the bit-exact oracle, and `-t N` through its 0MQ mux) against the same binary
with integration/libbwa_gpu_batch.so pre-loaded (hot path on the B200, one device call per phase and batch).
Also checks that the batched run's BAM records equal the `-t 1` run's, every record of the CPU sample.  The CPU arms run
oracle/_ref/bwa (the unmodified reference binary), the GPU arms integration/_host/bwa_host (the product's host: the
unmodified reference as a shared library) with the batched drivers pre-loaded.

    python scripts/bam2bam_bench.py --mode se --reads 2000000 --len 76 --genome-bp 100000000 [--cpu-sample 200000]
    python scripts/bam2bam_bench.py --mode pe --reads 1000000 --len 100 --genome-bp 100000000

The CPU arms run on a bounded prefix (--cpu-sample records) so the script ends within minutes; reads/s is per arm.
Index files are written by this repo's builder (byte-identical to `bwa index -a is`, tests/test_index.py)."""
import argparse
import importlib
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bamio  # noqa: E402

bwa = importlib.import_module("network-aware-bwa_b200")
REF_BWA = os.path.join(ROOT, "oracle", "_ref", "bwa")
DRIVER = os.path.join(ROOT, "integration", "_host", "bwa_host")
SHIM = os.path.join(ROOT, "integration", "libbwa_gpu_batch.so")


def write_bns(prefix, n, n_contigs):
    """.ann / .amb as bns_dump writes them (bntseq.c:63-85) for N-free contigs cut like simulate.write_fasta."""
    bounds = [n * i // n_contigs for i in range(n_contigs + 1)]
    with open(prefix + ".ann", "w") as f:
        f.write(f"{n} {n_contigs} 11\n")
        for c in range(n_contigs):
            f.write(f"0 chr{c + 1} (null)\n{bounds[c]} {bounds[c + 1] - bounds[c]} 0\n")
    with open(prefix + ".amb", "w") as f:
        f.write(f"{n} {n_contigs} 0\n")


def subset(reads, n):
    return bwa.simulate.Reads(reads.bases[: reads.offs[n]], reads.offs[: n + 1], reads.pos[:n], reads.strand[:n])


def run(prefix, bam_in, bam_out, threads, preload, extra, env_extra=None):
    env = dict(os.environ)
    if preload:
        env["LD_PRELOAD"] = SHIM
    env.update(env_extra or {})
    t0 = time.time()
    r = subprocess.run([DRIVER if preload else REF_BWA, "bam2bam", "-g", prefix, "-t", str(threads), *extra, "-f", bam_out, bam_in], capture_output=True, text=True, env=env)
    dt = time.time() - t0
    if r.returncode != 0:
        raise RuntimeError(r.stderr[-3000:])
    load = [l for l in r.stderr.splitlines() if "loading index" in l]
    load_s = float(load[0].split("...")[1].split()[0]) if load else 0.0
    return dt, load_s, r.stderr


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", choices=["se", "pe"], default="se")
    ap.add_argument("--reads", type=int, default=1_000_000, help="records (reads or pairs) in the GPU arm")
    ap.add_argument("--cpu-sample", type=int, default=100_000, help="records in the CPU arms")
    ap.add_argument("--len", type=int, default=76)
    ap.add_argument("--genome-bp", type=int, default=100_000_000)
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 8)
    ap.add_argument("--extra", default="", help="extra bam2bam options, e.g. '-l 1024 -n 0.01 -o 2'")
    ap.add_argument("--out", default="")
    ap.add_argument("--device", default="cuda:0")
    ap.add_argument("--gpu-env", action="append", default=[], help="a further run of the big GPU arm with these settings, e.g. name:BWAGPU_HOST_INFLATE=1,BWAGPU_TRACE=1")
    ap.add_argument("--log-dir", default="", help="keep every arm's whole stderr here")
    ap.add_argument("--stub", action="store_true", help="dry run without a GPU: tests/cpu_stub answers the device calls")
    a = ap.parse_args()
    dev = a.device
    if a.stub:
        global SHIM
        SHIM = os.path.join(ROOT, "tests", "cpu_stub", "libbwagpu_cpu_stub.so") + ":" + SHIM
    d = tempfile.mkdtemp(prefix="b2b_")
    t0 = time.time()
    T = bwa.simulate.make_genome(a.genome_bp, seed=1, repeat_frac=0.01)
    idx = bwa.index.build_index(T, device=dev)
    prefix = os.path.join(d, "g")
    bwa.index.save_index(prefix, idx)
    write_bns(prefix, a.genome_bp, 4)
    del idx
    print(f"[b2b] genome + index files {time.time() - t0:.1f}s", file=sys.stderr)
    if a.mode == "se":
        r1 = bwa.simulate.simulate_reads(T, a.reads, a.len, seed=1000, device=dev)
        r2 = None
    else:
        r1, r2 = bwa.simulate.simulate_pairs(T, a.reads, a.len, seed=1000, device=dev)
    big, small = os.path.join(d, "in.bam"), os.path.join(d, "in_small.bam")
    n_small = min(a.cpu_sample, a.reads)
    t0 = time.time()
    # BGZF-framed input, as sequencer pipelines / samtools write it (bamio.write_unaligned_bam's plain gzip stream would send
    # the shim down its one-thread reader: "input is not a BGZF file")
    bamio.write_unaligned_bam_fast(small, subset(r1, n_small), subset(r2, n_small) if r2 is not None else None)
    bamio.write_unaligned_bam_fast(big, r1, r2)
    print(f"[b2b] input BAMs {time.time() - t0:.1f}s", file=sys.stderr)
    extra = a.extra.split()
    per = 2 if a.mode == "pe" else 1
    res = {"mode": a.mode, "read_len": a.len, "genome_bp": a.genome_bp, "records_gpu": a.reads, "records_cpu": n_small, "host_threads": a.threads}

    def arm(name, bam, n, threads, preload, env_extra=None):
        out = os.path.join(d, name + ".bam")
        dt, load_s, log = run(prefix, bam, out, threads, preload, extra, env_extra)
        if a.log_dir:
            os.makedirs(a.log_dir, exist_ok=True)
            open(os.path.join(a.log_dir, f"b2b_{a.mode}_{name}.log"), "w").write(log)
        res[name] = {"wall_s": round(dt, 2), "index_load_s": load_s, "reads_per_s": round(n * per / max(dt - load_s, 1e-9)),
                     "log_tail": [l for l in log.splitlines() if ("processed in" in l and "(" in l) or "device calls" in l or "finish =" in l][-5:]}
        print(f"[b2b] {name}: {dt:.1f}s ({load_s:.1f}s index load) -> {res[name]['reads_per_s']} reads/s", file=sys.stderr)
        return out

    o_cpu1 = arm("cpu_t1", small, n_small, 1, False)
    arm("cpu_tN", small, n_small, a.threads, False)
    o_gpu_s = arm("gpu_batched_small", small, n_small, 1, True)
    arm("gpu_batched", big, a.reads, 1, True)
    for spec in a.gpu_env:
        name, _, kv = spec.partition(":")
        arm("gpu_batched_" + name, big, a.reads, 1, True, dict(x.split("=", 1) for x in kv.split(",") if x))
    x, y = bamio.read_bam_records(o_cpu1), bamio.read_bam_records(o_gpu_s)
    res["records_compared"] = len(x)
    res["records_differing"] = sum(1 for p, q in zip(x, y) if p != q) + abs(len(x) - len(y))
    line = json.dumps(res)
    print(line)
    if a.out:
        open(a.out, "w").write(line + "\n")


if __name__ == "__main__":
    main()
