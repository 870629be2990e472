#!/bin/bash
# round 2, GPU call 16: pipeline stages split, two destroy threads: shim parity tests + bench
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_inprocess_host.py tests/test_batched_bam2bam.py tests/test_gpu_parity.py -m gpu -x -q -k "inprocess or gpu_ or chunking" > $O/r2c16_host.log 2>&1; echo "host rc=$?" > $O/r2c16_box.log
timeout 1200 python bench.py --steps 5 --warmup 2 --no-cpu-baseline --no-aln-only > $O/r2c16_bench.json 2> $O/r2c16_bench.err
echo "bench rc=$?" >> $O/r2c16_box.log
cp /tmp/bench_host_rank0.log $O/r2c16_bench_host.log 2>/dev/null
tail -3 $O/r2c16_host.log
grep -E "pipelined|host CPU" $O/r2c16_bench_host.log | head -21 | tail -3
cat $O/r2c16_box.log
