#!/usr/bin/env python
"""bench.py -- reads/sec of the `bwa aln` hot path (K2 width + K3 gapped search + ordered
compaction) on BASELINE.json configs[1]: single-end 76 bp reads, default options
(-n 0.04 -o 1), vs a synthetic 100 Mb genome, one B200 per rank.

    python bench.py --gpus 1 --steps 3 --warmup 3            # this repo's CUDA path
    python bench.py --impl reference --steps 1 --warmup 0    # the reference's CPU path, all host cores
    torchrun ... bench.py --gpus N ...                       # weak scaling: one replica + one read shard per GPU

One JSON line on stdout (rank 0).  `value` = whole-job reads/s with the batch resident in
HBM (CUDA events inside the library, on the stream the kernels run on); `e2e` = the same
through the C-ABI call with host buffers (H2D and D2H inside the timed region); `roofline`
= occ-lookup bytes of the dominant kernel (k_search) against the measured HBM peak;
`cpu_baseline` = the reference's own bwa_cal_sa_reg_gap on this box's host cores on a
bounded sample.  Only the cpu_baseline / --impl reference legs and the parity spot-check
touch oracle/.
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

GENOME_BP = 100_000_000
READ_LEN = 76
READS_TOTAL = 10_000_000  # the configuration the metric is quoted on
METRIC = "reads/sec (aln: bwa_cal_sa_reg_gap per read = width bounds + gapped FM-index search)"
WORKLOAD = "SE 10M x 76bp, -n 0.04 -o 1, synthetic 100 Mb genome (BASELINE.json configs[1])"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


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


# ------------------------------------------------------------------ workload
def make_workload(bwa, n_reads: int, device: str, seed: int, genome_bp: int, read_len: int = READ_LEN):
    t0 = time.time()
    if genome_bp >= 1_000_000_000:  # cached genome + index files (workload.py)
        T, idx = bwa.workload.genome_and_index(genome_bp, seed=1, device=torch_device_index(device))
        t1 = t2 = time.time()
    else:
        T = bwa.simulate.make_genome(genome_bp, seed=1, repeat_frac=0.01)
        t1 = time.time()
        idx = bwa.index.build_index(T, device=device)
        t2 = time.time()
    reads = bwa.simulate.simulate_reads(T, n_reads, read_len, seed=seed, device=device)
    t3 = time.time()
    log(f"[bench] genome {t1 - t0:.1f}s, index ({device}) {t2 - t1:.1f}s, {n_reads} reads {t3 - t2:.1f}s")
    return T, idx, reads


def torch_device_index(device) -> int:
    d = str(device)
    return int(d.split(":")[1]) if ":" in d else 0


def seq_struct_array(abi, reads):
    """bwa_seq_t[n] over numpy storage, the way bam1_to_seq fills it (bwaseqio.c:272-297),
    built vectorised: returns (ctypes array pointer, keepalive)."""
    n = reads.n
    lens = reads.lens().astype(np.int64)
    assert (lens == lens[0]).all(), "vectorised builder expects fixed-length reads"
    L = int(lens[0])
    fwd = reads.bases.reshape(n, L)
    seq = np.ascontiguousarray(fwd[:, ::-1])
    rseq = np.where(seq > 3, 4, 3 - seq).astype(np.uint8)
    rec = np.zeros((n, 200), dtype=np.uint8)
    u64 = rec.view(np.uint64).reshape(n, 25)
    u32 = rec.view(np.uint32).reshape(n, 50)
    u64[:, 1] = seq.ctypes.data + np.arange(n, dtype=np.uint64) * np.uint64(L)    # seq   @ 8
    u64[:, 2] = rseq.ctypes.data + np.arange(n, dtype=np.uint64) * np.uint64(L)   # rseq  @ 16
    u32[:, 8] = L                                                                 # len:20 @ 32
    u32[:, 11] = L                                                                # clip_len @ 44
    u32[:, 28] = 0xFFFFFFFF                                                       # tid = -1 @ 112
    u32[:, 45] = L                                                                # full_len:20 @ 180
    ptr = C.cast(rec.ctypes.data, C.POINTER(abi.bwa_seq_t))
    return ptr, (rec, seq, rseq)


# ------------------------------------------------------------------ reference arm / cpu baseline
def time_reference(R, idx, reads, opt, target_s: float, threads: int):
    """Times the reference's own bwa_cal_sa_reg_gap (n_seqs = 1 per read, as bam2bam calls it)
    over `threads` host threads on a bounded prefix of the workload."""
    abi = R.abi
    ridx = R.RefIndex(idx)
    _, H = R.ref()
    probe = min(reads.n, 20000)

    def run(lo, hi):
        sub = R.bwa.simulate.Reads(reads.bases[reads.offs[lo]:reads.offs[hi]], reads.offs[lo:hi + 1] - reads.offs[lo], None, None)
        ptr, keep = seq_struct_array(abi, sub)
        t = time.perf_counter()
        H.refh_aln_batch(ridx.arr, hi - lo, ptr, C.byref(opt), threads)
        dt = time.perf_counter() - t
        H.refh_free_alns.argtypes = [C.c_int, C.POINTER(abi.bwa_seq_t)]
        H.refh_free_alns(hi - lo, ptr)
        return dt

    dt = run(0, probe)
    rate = probe / dt
    n = int(min(reads.n, max(probe, rate * target_s)))
    dt = run(0, n)
    return n / dt, n, dt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--reads", type=int, default=0, help="reads per step per GPU (default: the full 10M-read workload; reference arm: bounded sample)")
    ap.add_argument("--genome-bp", type=int, default=GENOME_BP)
    ap.add_argument("--read-len", type=int, default=READ_LEN, help="read length (default 76: configs[1]; 100 = the paired-end shape of configs[2]/[3])")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    host_cores = os.cpu_count() or 1

    import torch

    bwa = importlib.import_module("network-aware-bwa_b200")
    abi, api = bwa.abi, bwa.api
    opt = abi.default_gap_opt()  # -n 0.04 -o 1 are the defaults (bwtaln.c:19-35)

    # ------------------------------------------------------------- reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        import refload as R
        if not R.have_ref():
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref was not built (no /root/reference at build time)"}))
            return 0
        dev = "cuda" if torch.cuda.is_available() else "cpu"
        n_pool = args.reads or 400_000
        T, idx, reads = make_workload(bwa, n_pool, dev, seed=1000, genome_bp=args.genome_bp, read_len=args.read_len)
        target = 20.0
        rates, ns, total_t = [], [], 0.0
        for s in range(args.warmup + args.steps):
            rate, n, dt = time_reference(R, idx, reads, opt, target_s=target, threads=host_cores)
            if s >= args.warmup:
                rates.append(rate); ns.append(n); total_t += dt
        value = sum(ns) / total_t
        line = {
            "impl": "reference", "metric": METRIC, "value": value, "unit": "reads/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_t / max(1, args.steps),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample": f"{ns[0]} reads per step", "threads": host_cores},
            "cpu_baseline": {"value": value, "unit": "reads/s", "cores": host_cores, "kind": "reference",
                             "sample": f"{ns[0]}-read prefix of the workload per step, refh_aln_batch over {host_cores} threads"},
            "e2e": {"value": value, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------- B200 arm
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n_reads = args.reads or READS_TOTAL
    T, idx, reads = make_workload(bwa, n_reads, f"cuda:{local_rank}", seed=1000 + rank, genome_bp=args.genome_bp, read_len=args.read_len)
    torch.cuda.empty_cache()
    api.init([local_rank])
    api.load_index(idx)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: batch resident in HBM, device-timed
    api.resident_stage(reads.bases, reads.offs, opt)
    api.set_stats(True)
    api.resident_run()  # instrumented pass (untimed): algorithmic fetch counts of this batch
    st_counts = api.get_stats()
    api.set_stats(False)
    for _ in range(args.warmup):
        api.resident_run()
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    t_wall0 = time.perf_counter()
    dev_ms, search_ms, width_ms, launches = 0.0, 0.0, 0.0, 0
    tier_ms = [0.0] * 4
    for _ in range(args.steps):
        dev_ms += api.resident_run()
        st = api.get_stats()
        search_ms += st["ms_search"]; width_ms += st["ms_width"]; launches += st["launches"]
        tier_ms = [a + b for a, b in zip(tier_ms, st["ms_tier"])]
    barrier()
    wall_s = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    n_over2, n_over3 = st["n_overflow_t2"], st["n_overflow_t3"]
    if dist is not None:
        t = torch.tensor([dev_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms_max = float(t.item())
    else:
        dev_ms_max = dev_ms
    value = world * n_reads * args.steps / (dev_ms_max / 1e3)

    # ---- parity spot-check of the timed path against the reference (untimed)
    parity = None
    if rank == 0:
        try:
            import refload as R
            if R.have_ref():
                got = api.resident_fetch(reads.n)
                m = min(reads.n, 20000)
                sub = bwa.simulate.Reads(reads.bases[: reads.offs[m]], reads.offs[: m + 1], None, None)
                want = R.ref_aln(R.RefIndex(idx), sub, opt, threads=host_cores)
                got_sub = (got[0][:m], got[1][:m], got[2][: m + 1], got[3][: got[2][m]])
                errs = R.compare_aln(want, got_sub, "bench")
                parity = {"reads": m, "mismatches": len(errs), "checker": "oracle/_ref bwa_cal_sa_reg_gap"}
        except Exception as e:  # the checker is optional on the bench box
            parity = {"error": str(e)[:200]}

    # ---- e2e: through the reference-facing C-ABI call, host buffers, copies inside the timed region
    e2e = None
    if not args.no_e2e:
        ptr, keep = seq_struct_array(abi, reads)
        lib = api.lib()
        for _ in range(min(args.warmup, 2)):
            assert lib.bwa_gpu_cal_sa_reads_gap(reads.n, ptr, C.byref(opt)) == 0, lib.bwa_gpu_last_error()
            lib.bwa_gpu_free_alns(reads.n, ptr)
        barrier()
        e2e_s = 0.0
        e2e_steps = []
        n_aln_tot = 0
        for _ in range(args.steps):
            t0 = time.perf_counter()
            rc = lib.bwa_gpu_cal_sa_reads_gap(reads.n, ptr, C.byref(opt))
            e2e_steps.append(1e3 * (time.perf_counter() - t0))
            e2e_s += e2e_steps[-1] / 1e3
            assert rc == 0, lib.bwa_gpu_last_error()
            n_aln_tot = int(api.get_stats()["n_aln"])
            lib.bwa_gpu_free_alns(reads.n, ptr)  # untimed: the caller's bwa_free_read_seq1
        barrier()
        if dist is not None:
            t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = float(t.item())
        e2e = {"value": world * n_reads * args.steps / e2e_s, "unit": "reads/s",
               "h2d_bytes_per_step": int(reads.bases.size + 16 * reads.n),
               "d2h_bytes_per_step": int(8 * reads.n + 16 * n_aln_tot),
               "api": "bwa_gpu_cal_sa_reads_gap(n, bwa_seq_t*, gap_opt_t*) incl. per-read calloc of aln[]",
               "ms_each_step_rank0": [round(x, 1) for x in e2e_steps]}

    # ---- cpu baseline (rank 0, N = 1 only): the reference itself on a bounded sample
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            import refload as R
            if R.have_ref():
                rate, n, dt = time_reference(R, idx, reads, opt, target_s=15.0, threads=host_cores)
                cpu_baseline = {"value": rate, "unit": "reads/s", "cores": host_cores, "kind": "reference",
                                "sample": f"first {n} reads of the step's batch, {dt:.1f} s, refh_aln_batch "
                                          f"(bwa_cal_sa_reg_gap per read) over {host_cores} threads"}
        except Exception as e:
            cpu_baseline = {"error": str(e)[:200]}

    # ---- roofline context (rank 0): what pure dependent random 32-byte sector gathers sustain on this device
    random_sector = None
    if rank == 0 and not args.no_extras:
        try:
            random_sector = {"buffer_bytes": 2_300_000_000, "unit": "GB/s",
                             "by_chains_per_thread": {str(ch): api.probe_random_sectors(2_300_000_000, ch, 256) for ch in (1, 2, 4, 8)},
                             "index_sized_buffer": {"buffer_bytes": int(args.genome_bp), "chains4": api.probe_random_sectors(int(args.genome_bp), 4, 256)},
                             "note": "k_probe_gather: 8 x 256 threads per SM, each chain's next index depends on the sector just loaded"}
        except Exception as e:
            random_sector = {"error": str(e)[:200]}

    # ---- secondary kernels of the path (rank 0, N = 1): K4 SA->coordinate and K5 mate-rescue SW
    extras = None
    if rank == 0 and world == 1 and not args.no_extras:
        rng = np.random.default_rng(5)
        nq = 4_000_000
        sa_k = rng.integers(1, idx.bwt[0].seq_len + 1, size=nq, dtype=np.uint32)
        which = rng.integers(0, 2, size=nq, dtype=np.uint8)
        api.cal_pac_pos(sa_k[:1000], which[:1000])
        t0 = time.perf_counter(); api.cal_pac_pos(sa_k, which); dt4 = time.perf_counter() - t0
        steps4 = float(np.mean(sa_k % 32))  # LF steps to the next sampled row when rows are uniform
        nj, wlen, rlen = 200_000, 380, 100  # 2x100 bp, sigma 30: window = 6 sigma + 2 len (bwape.c:531-543)
        begs = rng.integers(0, idx.l_pac - wlen - 1, size=nj)
        sw_jobs = (abi.sw_job_t * nj)()
        qkeep = []
        for j in range(nj):
            b = int(begs[j]); o = int(rng.integers(0, wlen - rlen))
            q = T[b + o:b + o + rlen].copy(); q[rng.integers(0, rlen, size=3)] ^= 1
            qkeep.append(q)
            sw_jobs[j].beg, sw_jobs[j].reglen, sw_jobs[j].len = b, wlen, rlen
            sw_jobs[j].seq = q.ctypes.data_as(C.POINTER(C.c_ubyte))
        sw_res = (abi.sw_res_t * nj)()
        lib = api.lib()
        assert lib.bwa_gpu_mate_sw(1000, sw_jobs, sw_res) == 0
        t0 = time.perf_counter(); rc = lib.bwa_gpu_mate_sw(nj, sw_jobs, sw_res); dt5 = time.perf_counter() - t0
        assert rc == 0, lib.bwa_gpu_last_error()
        sw_ms = api.get_stats()["ms_sw_kernel"]
        extras = {"k4_sa": {"queries_per_s": nq / dt4, "host_call_ms": dt4 * 1e3, "lf_steps_per_query": steps4,
                            "algorithmic_gb_s": 64.0 * steps4 * nq / dt4 / 1e9, "note": "host buffers in and out"},
                  "k5_sw": {"jobs_per_s": nj / dt5, "host_call_ms": dt5 * 1e3, "gcups_forward": nj * wlen * rlen / dt5 / 1e9,
                            "kernel_ms": sw_ms, "gcups_forward_kernel": nj * wlen * rlen / (sw_ms / 1e3) / 1e9,
                            "shape": f"{wlen} x {rlen}", "note": "host buffers in and out; forward + reverse pass"}}

        # C5 of BASELINE.json (a parity-test configuration, reported here as a secondary number): ancient-DNA-style 30-50 bp
        # reads, seeding off (-l 1024), -n 0.01 -o 2.  ~10 % of the reads outgrow pass 0 and go through the warp-per-read
        # pass (csrc/search_warp.cuh); the batch is resident, device-timed like `value`.
        try:
            import refload as R
            n5 = 1_000_000
            reads5 = bwa.simulate.simulate_reads(T, n5, (30, 50), seed=1000, device=f"cuda:{local_rank}", adna=True, sub_rate=0.01)
            opt5 = abi.default_gap_opt(seed_len=1024, fnr=0.01, max_gapo=2)
            api.resident_stage(reads5.bases, reads5.offs, opt5)
            api.resident_run()
            ms5 = api.resident_run(); st5 = api.get_stats()
            c5 = {"workload": "1M aDNA-style 30-50bp reads, -l 1024 -n 0.01 -o 2, same 100 Mb genome (BASELINE.json configs[4])",
                  "reads_per_s": n5 / (ms5 / 1e3), "ms": ms5, "pass_ms": st5["ms_tier"][:3], "width_ms": st5["ms_width"],
                  "reads_through_warp_pass": int(st5["n_overflow_t2"]), "reads_through_guaranteed_pass": int(st5["n_overflow_t3"])}
            if R.have_ref():
                got5 = api.resident_fetch(n5)
                m5 = 100_000
                sub5 = bwa.simulate.Reads(reads5.bases[: reads5.offs[m5]], reads5.offs[: m5 + 1], None, None)
                want5 = R.ref_aln(R.RefIndex(idx), sub5, opt5, threads=host_cores); dt5r = R.ref_aln.last_batch_s
                got5s = (got5[0][:m5], got5[1][:m5], got5[2][: m5 + 1], got5[3][: got5[2][m5]])
                c5["cpu_reference_reads_per_s"] = m5 / dt5r
                c5["cpu_sample"] = f"first {m5} reads, {dt5r:.1f} s, bwa_cal_sa_reg_gap per read over {host_cores} threads"
                c5["parity_mismatches"] = len(R.compare_aln(want5, got5s, "c5"))
            extras["c5_adna"] = c5
        except Exception as e:
            extras["c5_adna"] = {"error": str(e)[:200]}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs, copy)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md 6.65 TB/s)"
    fetches = st_counts["occ_fetches_search"]  # reference-layout block fetches of ONE pass over the batch
    achieved = 64.0 * fetches / (search_ms / args.steps / 1e3) / 1e9
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "k_search_traffic.json"))).get("dram_bytes_per_launch")
    except Exception:
        pass
    line = {
        "metric": METRIC,
        "value": value, "unit": "reads/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32", "data": "synthetic",
        "config": {"workload": WORKLOAD if (args.read_len, args.genome_bp) == (READ_LEN, GENOME_BP) else
                   f"SE {n_reads} x {args.read_len}bp, -n 0.04 -o 1, synthetic {args.genome_bp} bp genome (non-default shape)",
                   "reads_per_step_per_gpu": n_reads, "read_len": args.read_len, "genome_bp": args.genome_bp,
                   "parallelism": f"replica x{world}, reads sharded, no collective",
                   "l2": "inputs larger than L2 (index 100 MB + width arena and search stacks of several GB per step); same batch every step",
                   "timed": "CUDA events on the library stream around K2+K3(+tiers)+compaction",
                   "tier2_reads": int(n_over2), "tier3_reads": int(n_over3)},
        "e2e": e2e,
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "kernel": "k_search (all tiers)", "peak_source": peak_src,
                     "algorithmic_bytes": "64 B x occ-block fetches of the reference layout (SURVEY.md §8d)",
                     "fetches_per_read": fetches / n_reads, "own_32B_blocks_per_read": st_counts["own_fetches_search"] / n_reads,
                     "kernel_ms_per_step": search_ms / args.steps, "width_ms_per_step": width_ms / args.steps,
                     "tier_ms_per_step": [t / args.steps for t in tier_ms],
                     "pops_per_read": st_counts["n_pops"] / n_reads, "pushes_per_read": st_counts["n_pushes"] / n_reads,
                     "stored_pushes_per_read": st_counts["n_stored"] / n_reads,
                     "per_read": {k: st_counts["n_" + k] / n_reads for k in ("pruned", "expand", "exact", "derive", "trips")},
                     "stats_pass_ms": {"queue_empty": st_counts["ns_queue_empty"] / 1e6, "kernel": st_counts["ns_kernel"] / 1e6},
                     "own_sector_gb_s": 32.0 * st_counts["own_fetches_search"] / (search_ms / args.steps / 1e3) / 1e9,
                     "random_sector_probe": random_sector},
        "cpu_baseline": cpu_baseline,
        "parity_sample": parity,
        "other_kernels": extras,
        "wall_s_timed_region": wall_s,
    }
    print(json.dumps(line))
    api.destroy()
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
