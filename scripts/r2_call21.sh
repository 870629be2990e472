#!/bin/bash
# round 2, GPU call 21: full GPU suite + bench (K = 5) with the device inflate on by default
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2c21_pytest.log 2>&1; echo "pytest rc=$?" > $O/r2c21_box.log
timeout 1200 python bench.py --steps 5 --warmup 2 > $O/r2c21_bench.json 2> $O/r2c21_bench.err
echo "bench rc=$?" >> $O/r2c21_box.log
cp /tmp/bench_host_rank0.log $O/r2c21_bench_host.log 2>/dev/null
tail -3 $O/r2c21_pytest.log; cat $O/r2c21_box.log
