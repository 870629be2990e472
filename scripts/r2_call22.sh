#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python bench.py --steps 8 --warmup 2 --no-cpu-baseline --no-aln-only --no-parity > $O/r2c22_bench.json 2> $O/r2c22_bench.err
echo "bench rc=$?"
cp /tmp/bench_host_rank0.log $O/r2c22_bench_host.log 2>/dev/null
BWAGPU_HOST_INFLATE=1 timeout 1200 python bench.py --steps 8 --warmup 2 --no-cpu-baseline --no-aln-only --no-parity > $O/r2c22_bench_hostinf.json 2> $O/r2c22_bench_hostinf.err
echo "bench hostinf rc=$?"
