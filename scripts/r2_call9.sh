#!/bin/bash
# round 2, GPU call 9: lane groups + service lane + two align threads; bench with the device-mode value
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2c9_pytest.log 2>&1; echo "pytest rc=$?" > $O/r2c9_box.log
timeout 1200 python bench.py --steps 3 --warmup 2 > $O/r2c9_bench.json 2> $O/r2c9_bench.err
echo "bench rc=$?" >> $O/r2c9_box.log
cp /tmp/bench_host_rank0.log $O/r2c9_bench_host.log 2>/dev/null
BWAGPU_CALL_GROUPS=1 BWAGPU_LANES=3 timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-aln-only --no-parity > $O/r2c9_bench_g1.json 2> $O/r2c9_bench_g1.err
echo "bench g1 rc=$?" >> $O/r2c9_box.log
tail -5 $O/r2c9_pytest.log
grep -E "pipelined" $O/r2c9_bench_host.log | head -12 | tail -4
cat $O/r2c9_box.log
