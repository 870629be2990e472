"""Where does the end-to-end time of bwa_gpu_cal_sa_reads_gap / bwa_gpu_aln_flat go?"""
import importlib, sys, os, time, ctypes as C, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
bwa = importlib.import_module("network-aware-bwa_b200")
import bench
api, abi = bwa.api, bwa.abi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
T, idx, reads = bench.make_workload(bwa, n, "cuda:0", seed=1000, genome_bp=100_000_000)
opt = abi.default_gap_opt()
api.init([0]); api.load_index(idx)
lib = api.lib()
for rep in range(2):
    t = time.perf_counter(); r = api.aln_flat(reads.bases, reads.offs, opt); dt = time.perf_counter() - t
    st = api.get_stats()
    print("flat  : %.0f ms  %.2f M reads/s | marshal %.0f h2d %.0f width %.0f search %.0f compact %.0f d2h %.0f device_total %.0f" % (
        dt * 1e3, n / dt / 1e6, st["ms_host_marshal"], st["ms_h2d"], st["ms_width"], st["ms_search"], st["ms_compact"], st["ms_d2h"], st["ms_total_device"]))
ptr, keep = bench.seq_struct_array(abi, reads)
for rep in range(2):
    t = time.perf_counter(); rc = lib.bwa_gpu_cal_sa_reads_gap(n, ptr, C.byref(opt)); dt = time.perf_counter() - t
    st = api.get_stats()
    print("struct: %.0f ms  %.2f M reads/s | marshal %.0f h2d %.0f width %.0f search %.0f compact %.0f d2h %.0f device_total %.0f" % (
        dt * 1e3, n / dt / 1e6, st["ms_host_marshal"], st["ms_h2d"], st["ms_width"], st["ms_search"], st["ms_compact"], st["ms_d2h"], st["ms_total_device"]))
    t = time.perf_counter(); lib.bwa_gpu_free_alns(n, ptr); print("  free() of the aln arrays (untimed in bench): %.0f ms" % ((time.perf_counter() - t) * 1e3))
api.destroy()
