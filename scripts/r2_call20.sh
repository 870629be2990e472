#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
BWAGPU_DEVICE_INFLATE=1 timeout 1200 python bench.py --steps 5 --warmup 2 --no-cpu-baseline --no-aln-only > $O/r2c20_bench_devinf.json 2> $O/r2c20_bench_devinf.err
echo "bench devinf rc=$?"
cp /tmp/bench_host_rank0.log $O/r2c20_bench_host.log 2>/dev/null
grep -E "pipelined|host CPU" $O/r2c20_bench_host.log | head -21 | tail -3
