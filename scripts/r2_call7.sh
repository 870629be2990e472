#!/bin/bash
# round 2, GPU call 7: K6 warp kernel parity, whole GPU suite, bench with the new defaults
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sw or global" > $O/r2c7_k6.log 2>&1; echo "k6 rc=$?" > $O/r2c7_box.log
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2c7_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2c7_box.log
timeout 900 python bench.py --steps 3 --warmup 2 > $O/r2c7_bench.json 2> $O/r2c7_bench.err
echo "bench rc=$?" >> $O/r2c7_box.log
cp /tmp/bench_host_rank0.log $O/r2c7_bench_host.log 2>/dev/null
tail -5 $O/r2c7_k6.log; tail -5 $O/r2c7_pytest.log
grep -E "host CPU|pipelined|device calls" $O/r2c7_bench_host.log | tail -4
cat $O/r2c7_box.log
