#!/usr/bin/env python
"""Device BGZF inflate alone: kernel time and throughput on members written by zlib (level 1) from a BAM-shaped stream.
TEST/BENCH INFRASTRUCTURE.    python scripts/inflate_bench.py [records]"""
import importlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
bwa = importlib.import_module("network-aware-bwa_b200")
from test_bgzf import bam_like, bgzf_file
api = bwa.api
n = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000
data = bam_like(20_000) * (n // 20_000)
packed = bgzf_file(data, 1)
api.init([0])
api.bgzf_inflate(packed)
best = None
for _ in range(3):
    t0 = time.perf_counter(); got, ooff, ms = api.bgzf_inflate(packed); dt = time.perf_counter() - t0
    best = ms if best is None else min(best, ms)
assert got == data
print(json.dumps({"members": int(len(ooff) - 1), "bytes_in": len(packed), "bytes_out": len(data), "kernel_ms": best, "gb_per_s_out": len(data) / best / 1e6,
                  "host_call_ms": dt * 1e3, "round_trip": "ok"}))
api.destroy()
