#!/bin/bash
# round 2, GPU call 32: the final tree -- bench (K = 3) both arms as a last check of the JSON lines
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python bench.py --steps 3 --warmup 3 > $O/r2c32_bench.json 2> $O/r2c32_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/r2c32_bench_ref.json 2> $O/r2c32_bench_ref.err; echo "ref rc=$?"
