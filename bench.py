#!/usr/bin/env python
"""bench.py -- reads/sec of the whole alignment workflow (aln + sampe: `bwa bam2bam`, BAM in -> BAM out, both passes)
on BASELINE.json configs[3]: paired-end 2 x 100 bp reads against a synthetic 3.1 Gb genome, default options, one B200
per rank, against the reference's own `bwa bam2bam -t <host cores>` on the same box.

    python bench.py --gpus 1 --steps 3 --warmup 3            # this repo: hot path on the B200 behind the reference's host code
    python bench.py --impl reference --steps 3 --warmup 1    # the unmodified reference, `bam2bam -t <nproc>`, same input
    torchrun ... bench.py --gpus N ...                       # weak scaling: one index replica + one shard of pairs per GPU

A step = one bam2bam run over the rank's shard (--pairs pairs, default 500 000): the reference's own host code
(integration/_host/libbwahost.so, built from /root/reference unmodified) run IN-PROCESS with the batched drivers of
integration/libbwa_gpu_batch.so in place of its two loops, so every alignment call lands in libbwagpu.so:
bwa_gpu_cal_sa_reads_gap (K2 widths + K3 gapped search), bwa_gpu_cal_pac_pos (K4), bwa_gpu_mate_sw_path (K5 + K6),
bwa_gpu_global_align_seqs (K6).  One JSON line on stdout (rank 0):

  value     reads/s with the inputs of every device call resident in HBM: the step's reads over the kernel-only device time
            of all its calls (CUDA events inside the library, on the streams the kernels run on)
  e2e       reads/s of the whole run through the reference-facing entry point (bwa_bam_to_bam) with host buffers: BAM
            inflate, record parsing, every H2D/D2H copy, pairing, BAM rewrite and deflate inside the timed region (the
            index load, which the reference reports separately and a long job pays once, is not)
  roofline  occ-lookup bytes of the dominant kernel (k_search) against the measured HBM peak, timed alone on a resident batch
  cpu_baseline / --impl reference: `oracle/_ref/bwa bam2bam -t <cores>` (the reference compiled here, unmodified), wall minus
            the index load it prints.
Only the cpu_baseline / --impl reference legs and the parity check run anything under oracle/.
"""
from __future__ import annotations

import argparse
import ctypes as C
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

GENOME_BP = 3_100_000_000
READ_LEN = 100
PAIRS_PER_STEP = 1_000_000
METRIC = "reads/sec (aln+sampe: bam2bam pass 1 + pass 2, BAM in -> BAM out)"
REF_BWA = os.path.join(ROOT, "oracle", "_ref", "bwa")
SHIM = os.path.join(ROOT, "integration", "libbwa_gpu_batch.so")


def log(*a):
    print("[bench]", *a, file=sys.stderr, flush=True)


def workload_name(genome_bp: int, read_len: int) -> str:
    tag = "BASELINE.json configs[3]" if (genome_bp, read_len) == (GENOME_BP, READ_LEN) else "non-default shape"
    return (f"PE 2x{read_len}bp bam2bam (sampe pairing + mate-rescue SW), defaults -n 0.04 -o 1, synthetic {genome_bp / 1e9:.2f} Gb genome "
            f"({tag})")


# ------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------ workload files
def pairs_bam(bwa, prefix: str, T, n_pairs: int, read_len: int, seed: int, device: str) -> str:
    """The rank's input: an unaligned paired BAM (BGZF) of n_pairs simulated pairs, cached beside the genome."""
    import bamio
    path = f"{prefix}.pe{read_len}_{n_pairs}_s{seed}.bam"
    if os.path.exists(path):
        return path
    t0 = time.time()
    r1, r2 = bwa.simulate.simulate_pairs(T(), n_pairs, read_len, seed=seed, device=device)
    tmp = path + f".tmp{os.getpid()}"
    bamio.write_unaligned_bam_fast(tmp, r1, r2, threads=min(16, os.cpu_count() or 4), qual_seed=seed)
    os.replace(tmp, path)
    log(f"{n_pairs} pairs simulated and written in {time.time() - t0:.1f}s -> {path}")
    return path


def prefix_bam(src: str, n_pairs: int) -> str:
    """The first n_pairs pairs of a cached input BAM as a file of their own (records of one size: cut the byte stream)."""
    import gzip
    import struct
    from concurrent.futures import ThreadPoolExecutor
    import bamio
    dst = f"{src[:-4]}.first{n_pairs}.bam"
    if os.path.exists(dst):
        return dst
    with gzip.open(src, "rb") as f:
        head = f.read(8)
        (l_text,) = struct.unpack_from("<i", head, 4)
        head += f.read(l_text + 4)  # the text and n_ref (= 0: an unaligned BAM)
        first = f.read(4)
        (bs,) = struct.unpack("<i", first)
        body = first + f.read((4 + bs) * 2 * n_pairs - 4)
    data = head + body
    B = 65280
    with ThreadPoolExecutor(8) as ex:
        blocks = list(ex.map(lambda i: bamio._bgzf_block(data[i:i + B], 1), range(0, len(data), B)))
    tmp = dst + f".tmp{os.getpid()}"
    with open(tmp, "wb") as f:
        for b in blocks:
            f.write(b)
        f.write(bamio.BGZF_EOF)
    os.replace(tmp, dst)
    return dst


# ------------------------------------------------------------------ the reference: `bwa bam2bam -t N` as a process
def run_reference(prefix: str, bam_in: str, bam_out: str, threads: int):
    """-> (seconds of the run without the index load the reference prints, index load seconds)"""
    t0 = time.perf_counter()
    r = subprocess.run([REF_BWA, "bam2bam", "-g", prefix, "-t", str(threads), "-f", bam_out, bam_in], capture_output=True, text=True)
    dt = time.perf_counter() - t0
    if r.returncode != 0:
        raise RuntimeError("reference bam2bam failed:\n" + r.stderr[-2000:])
    load = [l for l in r.stderr.splitlines() if "loading index" in l]
    load_s = float(load[0].split("...")[1].split()[0]) if load else 0.0
    return dt - load_s, load_s


# ------------------------------------------------------------------ this repo: bam2bam in-process behind the batched drivers
class Report(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("wall_s", "index_load_s", "device_init_s", "pass1_s", "pass2_s", "dev_aln_s", "dev_sa_s",
                                          "dev_sw_s", "dev_ga_s", "inflate_cpu_s", "process_cpu_s")] + \
               [(n, C.c_int64) for n in ("calls_aln", "reads_aln", "calls_sa", "q_sa", "calls_sw", "jobs_sw", "calls_ga", "jobs_ga", "sequences")] + \
               [("dev_bgzf_s", C.c_double), ("calls_bgzf", C.c_int64), ("bytes_bgzf", C.c_int64)]

    def asdict(self):
        return {f[0]: getattr(self, f[0]) for f in self._fields_}


class Host:
    """integration/libbwa_gpu_batch.so loaded into this process; run() = the reference's bwa_bam_to_bam entry point."""

    def __init__(self):
        if not os.path.exists(SHIM):
            raise SystemExit(f"bench.py: {SHIM} is not built (python -c 'import __graft_entry__ as g; g.build()')")
        self.H = C.CDLL(SHIM)
        self.H.bwa_bam_to_bam.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.c_char_p]
        self.H.bwa_gpu_batch_last_report.argtypes = [C.POINTER(Report)]
        assert self.H.bwa_gpu_batch_report_size() == C.sizeof(Report), "bench.py's Report is out of step with integration/bwa_gpu_batch.h"
        self.H.bwa_gpu_batch_keep_index(1)

    def run(self, prefix: str, bam_in: str, bam_out: str):
        args = [b"bam2bam", b"-g", prefix.encode(), b"-t", b"1", b"-f", bam_out.encode(), bam_in.encode()]
        av = (C.c_char_p * (len(args) + 1))(*args, None)
        sys.stderr.flush()
        rc = self.H.bwa_bam_to_bam(len(args), av, b"bench")
        if rc != 0:
            raise RuntimeError(f"bwa_bam_to_bam returned {rc}")
        rep = Report()
        self.H.bwa_gpu_batch_last_report(C.byref(rep))
        return rep.asdict()

    def close(self):
        self.H.bwa_gpu_batch_drop_index()


def quiet_stderr(rank: int):
    """The reference chats on stderr (a line per batch and per read length); it goes to a file, the bench log stays readable."""
    if os.environ.get("BENCH_VERBOSE"):
        return None
    saved = os.dup(2)
    fd = os.open(os.path.join(os.environ.get("TMPDIR", "/tmp"), f"bench_host_rank{rank}.log"), os.O_WRONLY | os.O_CREAT | os.O_APPEND, 0o644)
    os.dup2(fd, 2)
    os.close(fd)
    return saved


def restore_stderr(saved):
    if saved is not None:
        sys.stderr.flush()
        os.dup2(saved, 2)
        os.close(saved)


class Outputs:
    """A fresh output path per run, as a user's successive jobs have; finished outputs are deleted by a background thread.
    (Writing every step over the previous step's file puts the deletion of that file's 250 MB -- open(O_TRUNC) at the start,
    ext4's flush-on-close of a truncated-and-rewritten file at the end: 0.1 s of a 1.3 s step -- inside the job.)"""

    def __init__(self, stem: str):
        import queue
        import threading
        self.stem, self.n, self.live = stem, 0, []
        self.q = queue.Queue()
        self.th = threading.Thread(target=self._reaper, daemon=True)
        self.th.start()

    def _reaper(self):
        while True:
            path = self.q.get()
            if path is None:
                return
            try:
                os.unlink(path)
            except OSError:
                pass

    def next(self) -> str:
        while len(self.live) > 1:  # keep the latest one (a reader may want it), drop the rest
            self.q.put(self.live.pop(0))
        path = f"{self.stem}.{os.getpid()}.{self.n}.bam"
        self.n += 1
        self.live.append(path)
        return path

    def close(self):
        for path in self.live:
            self.q.put(path)
        self.live = []
        self.q.put(None)
        self.th.join()


def records_differing(a_path: str, b_path: str):
    import bamio
    a, b = bamio.read_bam_records(a_path), bamio.read_bam_records(b_path)
    return len(a), sum(1 for x, y in zip(a, b) if x != y) + abs(len(a) - len(b))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pairs", type=int, default=PAIRS_PER_STEP, help="read pairs per step per GPU (both arms)")
    ap.add_argument("--genome-bp", type=int, default=GENOME_BP)
    ap.add_argument("--read-len", type=int, default=READ_LEN)
    ap.add_argument("--aln-reads", type=int, default=4_000_000, help="reads of the resident aln-only batch the k_search roofline is measured on")
    ap.add_argument("--cpu-sample-pairs", type=int, default=200_000)
    ap.add_argument("--ref-budget-s", type=float, default=200.0, help="--impl reference: seconds of reference CPU time the K timed steps may take in all")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-aln-only", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--parity-pairs", type=int, default=50_000, help="pairs of the shard run through `bam2bam -t 1` as well (untimed), every record compared")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    host_cores = os.cpu_count() or 1
    config = {"workload": workload_name(args.genome_bp, args.read_len), "pairs_per_step_per_gpu": args.pairs, "read_len": args.read_len,
              "genome_bp": args.genome_bp,
              "l2": f"inputs larger than L2 (re-laid-out index {args.genome_bp / 1e9:.2f} GB per replica); the same shard every step",
              "outputs": "every step writes a fresh output BAM; older outputs are deleted by a background thread of the harness"}

    import torch
    bwa = importlib.import_module("network-aware-bwa_b200")
    have_gpu = torch.cuda.is_available()

    # ------------------------------------------------------------- reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        if not os.path.exists(REF_BWA):
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/bwa was not built (no /root/reference at build time)"}))
            return 0
        cached = os.path.exists(os.path.join(bwa.workload.cache_root(), bwa.workload.genome_key(args.genome_bp, 1), ".done"))
        if not have_gpu and not cached and args.genome_bp > 50_000_000:  # small genomes: the CPU harness builder will do
            print(json.dumps({"impl": "reference", "unavailable": "the workload's index files are not cached and there is no device to build them on"}))
            return 0
        prefix = bwa.workload.ensure_genome_files(args.genome_bp, 1, 0)  # set-up only: the files `bwa index` would write
        bam = pairs_bam(bwa, prefix, lambda: bwa.workload.load_genome(prefix), args.pairs, args.read_len, 1000, "cuda:0" if have_gpu else "cpu")
        small = prefix_bam(bam, min(args.pairs, 20_000))
        # a bounded sample of the step's shard per step, so that K steps end within a few minutes whatever K is: the
        # reference runs ~35 k pairs/s on 16 cores, its rate does not depend on the length of the run
        ref_pairs = min(args.pairs, max(50_000, int(args.ref_budget_s * 2200 * host_cores / max(1, args.steps))))
        sample = bam if ref_pairs == args.pairs else prefix_bam(bam, ref_pairs)
        outs = Outputs(bam[:-4] + ".ref_out")
        for _ in range(args.warmup):  # page cache / index files warm: short runs
            run_reference(prefix, small, outs.next(), host_cores)
        total_t, loads, steps_ms = 0.0, [], []
        for _ in range(args.steps):
            dt, load_s = run_reference(prefix, sample, outs.next(), host_cores)
            total_t += dt; loads.append(load_s); steps_ms.append(round(dt * 1e3, 1))
        outs.close()
        value = 2 * ref_pairs * args.steps / total_t
        line = {
            "impl": "reference", "metric": METRIC, "value": value, "unit": "reads/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * total_t / max(1, args.steps), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": value, "unit": "reads/s", "cores": host_cores, "kind": "reference",
                             "sample": f"the first {ref_pairs} pairs of the step's {args.pairs}-pair shard per step, `bwa bam2bam -t {host_cores}` (unmodified reference), wall minus the "
                                       f"index load it prints ({np.mean(loads):.1f} s per run); warm-up runs on a 20k-pair prefix"},
            "e2e": {"value": value, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "ms_each_step": steps_ms, "sample_pairs_per_step": ref_pairs,
        }
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------- B200 arm
    if not have_gpu:
        raise SystemExit("bench.py: no CUDA device; this path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    threads = max(2, host_cores // world)
    os.environ.setdefault("BWAGPU_DEVICE", str(local_rank))
    os.environ.setdefault("BWAGPU_SHIM_THREADS", str(threads))
    os.environ.setdefault("BWAGPU_HOST_THREADS", str(max(1, min(8, threads // 3))))
    os.environ.setdefault("BWAGPU_BATCH_RECORDS", str(1 << 17))

    t0 = time.time()
    prefix = bwa.workload.ensure_genome_files(args.genome_bp, 1, local_rank)
    genome = {}

    def T():
        if "T" not in genome:
            genome["T"] = bwa.workload.load_genome(prefix)
        return genome["T"]

    bam = pairs_bam(bwa, prefix, T, args.pairs, args.read_len, 1000 + rank, f"cuda:{local_rank}")
    outs = Outputs(f"{bam[:-4]}.out")  # a fresh output file per run
    aln_reads = None
    if rank == 0 and not args.no_aln_only:
        aln_reads = bwa.simulate.simulate_reads(T(), args.aln_reads, args.read_len, seed=1000, device=f"cuda:{local_rank}")
    genome.clear()
    torch.cuda.empty_cache()
    log(f"rank {rank}: workload ready in {time.time() - t0:.1f}s ({threads} host threads per rank)")

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    api = bwa.api
    host = Host()
    saved = quiet_stderr(rank)
    try:
        for _ in range(max(1, args.warmup)):  # the first run also loads the index and sets the device up
            host.run(prefix, bam, outs.next())
        api.reset_totals()
        sampler = ClockSampler(local_rank)
        sampler.start()
        barrier()
        t_wall0 = time.perf_counter()
        e2e_s, steps_ms, reps = 0.0, [], []
        for _ in range(args.steps):
            t1 = time.perf_counter()
            rep = host.run(prefix, bam, outs.next())
            dt = time.perf_counter() - t1 - rep["index_load_s"]
            e2e_s += dt; steps_ms.append(round(dt * 1e3, 1)); reps.append(rep)
        barrier()
        wall_s = time.perf_counter() - t_wall0
        tot_e2e = api.get_totals()
        # ---- the same K steps once more for `value`: the device time the job needs.  In the pipelined configuration above the
        # kernels of three lanes and of two passes' stages overlap, so their event times count the same device seconds
        # several times; here every pass is ONE batch on ONE lane -- kernels run one after the other, alone on the device --
        # and the sum of their event times is the device-busy time of the step.
        pipelined_env = {k: os.environ.get(k) for k in ("BWAGPU_LANES", "BWAGPU_BATCH_RECORDS", "BWAGPU_BATCH_RAMP", "BWAGPU_INFLATE_MEMBERS")}
        os.environ.update({"BWAGPU_LANES": "1", "BWAGPU_BATCH_RECORDS": str(args.pairs), "BWAGPU_BATCH_RAMP": "0", "BWAGPU_INFLATE_MEMBERS": "9216"})  # > the shard's ~8400 BGZF members
        host.H.bwa_gpu_batch_reset_device()
        host.run(prefix, bam, outs.next())  # untimed: sets the device up again, sizes the buffers
        api.reset_totals()
        barrier()
        t_dev0 = time.perf_counter()
        for _ in range(args.steps):
            host.run(prefix, bam, outs.next())
        barrier()
        wall_dev_s = time.perf_counter() - t_dev0
        clocks = sampler.stop()
        tot = api.get_totals()
        for k, v in pipelined_env.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
        host.H.bwa_gpu_batch_reset_device()
    finally:
        restore_stderr(saved)
        outs.close()
    kernel_ms = tot["ms_width"] + tot["ms_search"] + tot["ms_sa"] + tot["ms_sw"] + tot["ms_global"] + tot["ms_bgzf"] + tot["ms_inflate"]
    if dist is not None:
        t = torch.tensor([e2e_s, kernel_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s_max, kernel_ms_max = float(t[0].item()), float(t[1].item())
    else:
        e2e_s_max, kernel_ms_max = e2e_s, kernel_ms
    reads_per_step = 2 * args.pairs
    value = world * reads_per_step * args.steps / (kernel_ms_max / 1e3)
    e2e = {"value": world * reads_per_step * args.steps / e2e_s_max, "unit": "reads/s",
           "h2d_bytes_per_step": int(tot_e2e["h2d_bytes"] / args.steps), "d2h_bytes_per_step": int(tot_e2e["d2h_bytes"] / args.steps),
           "api": "bwa_bam_to_bam (the reference's bam2bam entry point, in-process) behind integration/libbwa_gpu_batch.so; BAM in -> BAM out, "
                  "both passes; wall minus the index load (0 after the first run: the index stays loaded, as in one long job)",
           "ms_each_step_rank0": steps_ms, "host_threads_per_rank": threads}
    last = reps[-1]
    pipeline = {k: last[k] for k in ("pass1_s", "pass2_s", "dev_aln_s", "dev_sa_s", "dev_sw_s", "dev_ga_s", "inflate_cpu_s", "process_cpu_s", "reads_aln", "q_sa",
                                     "jobs_sw", "jobs_ga", "dev_bgzf_s", "bytes_bgzf")}
    pipeline["device_call_share_of_wall"] = ((last["dev_aln_s"] + last["dev_sa_s"] + last["dev_sw_s"] + last["dev_ga_s"])
                                             / max(1e-9, last["wall_s"] - last["index_load_s"]))
    per_step = {k: tot[k] / args.steps for k in ("ms_width", "ms_search", "ms_sa", "ms_sw", "ms_global", "ms_bgzf", "ms_inflate")}
    per_step_pipelined = {k: tot_e2e[k] / args.steps for k in ("ms_width", "ms_search", "ms_sa", "ms_sw", "ms_global", "ms_bgzf", "ms_inflate")}

    if rank != 0:
        host.close()
        if dist is not None:
            dist.destroy_process_group()
        return 0

    # ---- parity (untimed): a prefix of the shard through both implementations, every record compared
    parity = None
    if not args.no_parity and os.path.exists(REF_BWA):
        try:
            n_par = min(args.pairs, args.parity_pairs)
            small = prefix_bam(bam, n_par)
            saved = quiet_stderr(rank)
            try:
                host.run(prefix, small, small[:-4] + ".gpu_out.bam")
            finally:
                restore_stderr(saved)
            run_reference(prefix, small, small[:-4] + ".ref_out.bam", 1)
            n, bad = records_differing(small[:-4] + ".ref_out.bam", small[:-4] + ".gpu_out.bam")
            parity = {"records": n, "records_differing": bad,
                      "checker": f"oracle/_ref/bwa bam2bam -t 1 on the first {n_par} pairs of the shard, whole records compared"}
        except Exception as e:  # the checker is optional on the bench box
            parity = {"error": str(e)[:300]}

    # ---- k_search alone on a resident batch: the roofline of the dominant kernel (and the aln-only throughput of round 1's line)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs, copy)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md 6.65 TB/s)"
    roofline, aln_only = None, None
    if aln_reads is not None:
        if parity is None or "error" in parity:
            # the value region ended with a device reset and no parity run has set the device (and its index) up again
            small = prefix_bam(bam, min(args.pairs, 20000))
            saved = quiet_stderr(rank)
            try:
                host.run(prefix, small, small[:-4] + ".gpu_out.bam")
            finally:
                restore_stderr(saved)
        opt = bwa.abi.default_gap_opt()
        api.resident_stage(aln_reads.bases, aln_reads.offs, opt)
        api.set_stats(True)
        api.resident_run()  # instrumented pass (untimed): algorithmic fetch counts of this batch
        st_counts = api.get_stats()
        api.set_stats(False)
        for _ in range(3):
            api.resident_run()
        dev_ms = search_ms = width_ms = 0.0
        n_rep = max(3, min(args.steps, 10))
        for _ in range(n_rep):
            dev_ms += api.resident_run()
            st = api.get_stats()
            search_ms += st["ms_search"]; width_ms += st["ms_width"]
        fetches = st_counts["occ_fetches_search"]
        achieved = 64.0 * fetches / (search_ms / n_rep / 1e3) / 1e9
        traffic = None
        try:
            # ncu's DRAM bytes of a 2 M-read launch of the same kernel on the same workload, scaled to this launch's reads
            traffic = json.load(open(os.path.join(ROOT, "profiles", "k_search_traffic.json")))["c4"]["dram_bytes_per_read"] * args.aln_reads
        except Exception:
            pass
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                    "kernel": "k_search (all passes), timed alone on a resident batch of the workload's reads",
                    "peak_source": peak_src, "algorithmic_bytes": "64 B x occ-block fetches of the reference layout (SURVEY.md §8d)",
                    "traffic_source": "profiles/k_search_traffic.json c4.dram_bytes_per_read (ncu, 2 M-read launch) x reads_per_launch",
                    "reads_per_launch": args.aln_reads, "fetches_per_read": fetches / args.aln_reads,
                    "own_32B_blocks_per_read": st_counts["own_fetches_search"] / args.aln_reads,
                    "kernel_ms_per_launch": search_ms / n_rep, "width_ms_per_launch": width_ms / n_rep,
                    "own_sector_gb_s": 32.0 * st_counts["own_fetches_search"] / (search_ms / n_rep / 1e3) / 1e9,
                    "per_read": {k: st_counts["n_" + k] / args.aln_reads for k in ("pops", "pushes", "stored", "pruned", "expand", "exact", "derive", "trips")}}
        aln_only = {"metric": "reads/sec (aln only: bwa_cal_sa_reg_gap per read = K2 + K3), resident batch, device-timed",
                    "value": args.aln_reads * n_rep / (dev_ms / 1e3), "reads": args.aln_reads, "read_len": args.read_len}
        try:
            roofline["random_sector_probe_gb_s"] = {"2.3GB_buffer_chains4": api.probe_random_sectors(2_300_000_000, 4, 256)}
        except Exception as e:
            roofline["random_sector_probe_gb_s"] = {"error": str(e)[:200]}
        try:
            # BASELINE.json's target for the occ lookup kernel: >= 50 % of the achievable random-sector throughput
            roofline["own_sector_frac_of_probe"] = roofline["own_sector_gb_s"] / roofline["random_sector_probe_gb_s"]["2.3GB_buffer_chains4"]
        except Exception:
            pass
        try:
            # what binds the kernel (profiles/r2_fetch_probe.md): the number of requests it sends to L2, not their bytes --
            # ncu's request counts per read x this launch's reads / its time, beside the request rate of the probe
            c4 = json.load(open(os.path.join(ROOT, "profiles", "k_search_traffic.json")))["c4"]
            rq = c4["l2_read_requests_per_read"] + c4["l2_write_requests_per_read"]
            roofline["l2_requests"] = {"per_read": rq, "g_per_s": rq * args.aln_reads / (search_ms / n_rep / 1e3) / 1e9,
                                       "probe_g_loads_per_s": roofline["random_sector_probe_gb_s"].get("2.3GB_buffer_chains4", 0.0) / 32.0,
                                       "source": c4["source"]}
        except Exception:
            pass

    # ---- the other kernels of the step, from the library's totals over the timed steps
    int_alu_peak = 148 * 128 * float(peaks.get("sm_max_mhz", 1965.0)) * 1e6  # integer lane-ops/s: 148 SMs x 128 lanes x clock

    def per_s(units, ms):
        return units / max(1e-9, ms / 1e3)

    other = {
        "k4_sa": {"queries_per_step": tot["sa_queries"] / args.steps, "kernel_ms_per_step": per_step["ms_sa"],
                  "queries_per_s_kernel": per_s(tot["sa_queries"], tot["ms_sa"]),
                  "algorithmic_gb_s": 64.0 * 31.0 * per_s(tot["sa_queries"], tot["ms_sa"]) / 1e9,
                  "frac_of_hbm_peak": 64.0 * 31.0 * per_s(tot["sa_queries"], tot["ms_sa"]) / 1e9 / peak,
                  "unit": "64 B x 31 LF steps per query (bwt.c:72-81: rows are sampled, k % sa_intv == 0 with sa_intv 32, so the walk length is geometric with mean 31; ncu counts 32.2 sector loads per query, profiles/r2_k_sa.md)"},
        "k5_sw": {"jobs_per_step": tot["sw_jobs"] / args.steps, "kernel_ms_per_step": per_step["ms_sw"],
                  "gcups_forward_kernel": per_s(tot["sw_cells_fwd"], tot["ms_sw"]) / 1e9,
                  "integer_alu_bound_gcups": int_alu_peak / 12 / 1e9,
                  "frac_of_integer_alu_bound": per_s(tot["sw_cells_fwd"], tot["ms_sw"]) / (int_alu_peak / 12),
                  "bound": "integer ALU: 148 SMs x 128 lanes x SM clock lane-ops/s at 12 lane-ops per affine-gap cell"},
        "k6_global": {"jobs_per_step": (tot["sw_jobs"] + tot["ga_jobs"]) / args.steps, "kernel_ms_per_step": per_step["ms_global"]},
        "k2_width": {"kernel_ms_per_step": per_step["ms_width"]},
        "bgzf_deflate": {"bytes_in_per_step": tot["bgzf_bytes_in"] / args.steps, "bytes_out_per_step": tot["bgzf_bytes_out"] / args.steps,
                         "kernel_ms_per_step": per_step["ms_bgzf"], "gb_per_s_in": per_s(tot["bgzf_bytes_in"], tot["ms_bgzf"]) / 1e9,
                         "ratio": tot["bgzf_bytes_out"] / max(1, tot["bgzf_bytes_in"]),
                         "bound": "latency of the per-block LZ77 / Huffman phases (one CTA per 64 KB block, shared memory only); not an HBM-bound kernel"},
        "bgzf_inflate": {"bytes_in_per_step": tot["inflate_bytes_in"] / args.steps, "bytes_out_per_step": tot["inflate_bytes_out"] / args.steps,
                         "kernel_ms_per_step": per_step["ms_inflate"], "gb_per_s_out": per_s(tot["inflate_bytes_out"], tot["ms_inflate"]) / 1e9,
                         "bound": "latency of lane 0's serial walk of each member's bit stream (one warp per member)"},
        "k3_search_in_job": {"kernel_ms_per_step": per_step["ms_search"], "pass_ms_per_step": [x / args.steps for x in tot["ms_search_pass"]]},
    }
    try:
        # K4 is a dependent random gather: 31 block loads + 1 SA read per query, against what the random-sector probe sustains
        probe_g = roofline["random_sector_probe_gb_s"]["2.3GB_buffer_chains4"] / 32.0
        other["k4_sa"]["requests_g_per_s"] = 32.0 * other["k4_sa"]["queries_per_s_kernel"] / 1e9
        other["k4_sa"]["frac_of_random_sector_probe"] = other["k4_sa"]["requests_g_per_s"] / probe_g
    except Exception:
        pass

    # ---- cpu baseline (N = 1 only): the reference itself on a bounded prefix of the shard
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline and os.path.exists(REF_BWA):
        try:
            n_s = min(args.pairs, args.cpu_sample_pairs)
            sample = prefix_bam(bam, n_s)
            dt, load_s = run_reference(prefix, sample, sample[:-4] + ".ref_out.bam", host_cores)
            cpu_baseline = {"value": 2 * n_s / dt, "unit": "reads/s", "cores": host_cores, "kind": "reference",
                            "sample": f"first {n_s} pairs of the step's shard, `oracle/_ref/bwa bam2bam -t {host_cores}` (unmodified reference), "
                                      f"{dt:.1f} s after the {load_s:.1f} s index load it prints"}
            if n_s >= 20_000:
                dt1, _ = run_reference(prefix, prefix_bam(bam, 20_000), sample[:-4] + ".ref_t1_out.bam", 1)
                cpu_baseline["one_thread_reads_per_s"] = 2 * 20_000 / dt1
        except Exception as e:
            cpu_baseline = {"error": str(e)[:300]}

    line = {
        "metric": METRIC, "value": value, "unit": "reads/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": kernel_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32", "data": "synthetic", "config": config,
        "e2e": e2e, "gpu_launches": int(tot["launches"] + tot_e2e["launches"]), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline,
        "parity_sample": parity, "aln_only": aln_only, "other_kernels": other, "pipeline_last_step_rank0": pipeline,
        "kernel_ms_per_step": per_step, "kernel_ms_per_step_pipelined_runs": per_step_pipelined, "parallelism": f"one process + one index replica + one shard of pairs per GPU (x{world}), no collective",
        "timed": "two timed regions of K steps each, both bracketed by barrier + synchronize.  e2e: perf_counter around bwa_bam_to_bam in the "
                 "pipelined configuration (4 lanes in 2 groups, batches of 131072 records).  value: the same K runs with one batch per pass on one "
                 "lane, so that no two kernels overlap; CUDA events inside the library around every kernel (K2, K3, K4, K5, K6, BGZF inflate and deflate), "
                 "summed = device-busy time of the job (kernel_ms_per_step); the pipelined runs' event sums are in "
                 "kernel_ms_per_step_pipelined_runs (overlapping kernels counted more than once)",
        "wall_s_timed_region": wall_s + wall_dev_s, "wall_s_e2e_region": wall_s, "wall_s_value_region": wall_dev_s,
    }
    print(json.dumps(line))
    host.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
