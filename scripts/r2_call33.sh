#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-aln-only --no-parity > $O/r2c33_bench.json 2> $O/r2c33_bench.err; echo "bench rc=$?"
cp /tmp/bench_host_rank0.log $O/r2c33_bench_host.log 2>/dev/null
grep -E "pipelined|host CPU" $O/r2c33_bench_host.log | sed -n 16,18p
