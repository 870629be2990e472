#!/usr/bin/env python
"""Device BGZF deflate alone: kernel time and throughput on a BAM-shaped stream (bwa_gpu_bgzf_deflate), every member
inflated with zlib and compared.  TEST/BENCH INFRASTRUCTURE.
    python scripts/bgzf_bench.py [records]"""
import importlib, json, os, sys, time, zlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
bwa = importlib.import_module("network-aware-bwa_b200")
from test_bgzf import bam_like, check_members
api = bwa.api
n = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000
data = bam_like(20_000) * (n // 20_000)
api.init([0])
api.bgzf_deflate(data[:1 << 20], 2)
best = None
for _ in range(3):
    t0 = time.perf_counter(); packed, lens, ms = api.bgzf_deflate(data, 2); dt = time.perf_counter() - t0
    best = ms if best is None else min(best, ms)
check_members(data, packed, lens)
z2 = sum(len(zlib.compress(data[i:i + 65280], 2)) for i in range(0, min(len(data), 20 * 65280), 65280))
mine = int(lens[:20].sum())
print(json.dumps({"bytes_in": len(data), "bytes_out": len(packed), "members": int(len(lens)), "kernel_ms": best, "gb_per_s_in": len(data) / best / 1e6,
                  "host_call_ms": dt * 1e3, "first_20_members_bytes": mine, "zlib_level2_same_20_blocks": z2 + 26 * 20, "round_trip": "ok"}))
api.destroy()
