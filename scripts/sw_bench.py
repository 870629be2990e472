"""K5 throughput: jobs/s and forward-pass cell updates/s of bwa_gpu_mate_sw (kernel time and host-call time)."""
import importlib, sys, os, time, ctypes as C, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
bwa = importlib.import_module("network-aware-bwa_b200")
api, abi = bwa.api, bwa.abi
T = bwa.simulate.make_genome(20_000_000, seed=1)
idx = bwa.index.build_index(T, device="cuda:0")
api.init([0]); api.load_index(idx)
lib = api.lib()
rng = np.random.default_rng(5)
for (wlen, rlen, nj) in ((380, 100, 400_000), (260, 76, 400_000), (800, 150, 100_000)):
    begs = rng.integers(0, idx.l_pac - wlen - 1, size=nj)
    jobs = (abi.sw_job_t * nj)(); keep = []
    for j in range(nj):
        b = int(begs[j]); o = int(rng.integers(0, wlen - rlen))
        q = T[b + o:b + o + rlen].copy(); q[rng.integers(0, rlen, size=3)] ^= 1
        keep.append(q); jobs[j].beg, jobs[j].reglen, jobs[j].len = b, wlen, rlen
        jobs[j].seq = q.ctypes.data_as(C.POINTER(C.c_ubyte))
    res = (abi.sw_res_t * nj)()
    lib.bwa_gpu_mate_sw(1000, jobs, res)
    t0 = time.perf_counter(); rc = lib.bwa_gpu_mate_sw(nj, jobs, res); dt = time.perf_counter() - t0
    ms = api.get_stats()["ms_sw_kernel"]
    print(f"{wlen}x{rlen}: K5 kernel {ms:.1f} ms = {nj / ms / 1e3:.2f} M jobs/s, {nj * wlen * rlen / ms / 1e6:.0f} GCUPS (forward cells); host call {dt * 1e3:.0f} ms")
    pres = (abi.path_res_t * nj)(); pool = C.c_void_p()
    t0 = time.perf_counter(); rc = lib.bwa_gpu_mate_sw_path(nj, jobs, pres, C.byref(pool)); dt = time.perf_counter() - t0
    ms2 = api.get_stats()["ms_sw_kernel"]
    print(f"    K5+K6 kernels {ms2:.1f} ms = {nj / ms2 / 1e3:.2f} M jobs/s; host call {dt * 1e3:.0f} ms")
    t0 = time.perf_counter(); rc = lib.bwa_gpu_global_align(nj, jobs, 5, 50, pres, C.byref(pool)); dt = time.perf_counter() - t0
    ms3 = api.get_stats()["ms_sw_kernel"]
    print(f"    K6 alone on the whole window (gap_end 5, band 50) {ms3:.1f} ms = {nj / ms3 / 1e3:.2f} M jobs/s; host call {dt * 1e3:.0f} ms")
api.destroy()
