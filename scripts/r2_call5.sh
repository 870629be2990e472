#!/bin/bash
# round 2, GPU call 5: device BGZF deflate -- parity tests, then the in-process bam2bam bench with it and with the zlib writer
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_bgzf.py -m gpu -x -q > $O/r2c5_bgzf.log 2>&1; echo "bgzf rc=$?" > $O/r2c5_box.log
timeout 900 python -m pytest tests/test_inprocess_host.py tests/test_batched_bam2bam.py -m gpu -x -q > $O/r2c5_host.log 2>&1; echo "host rc=$?" >> $O/r2c5_box.log
timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-aln-only --no-parity > $O/r2c5_bench.json 2> $O/r2c5_bench.err
echo "bench rc=$?" >> $O/r2c5_box.log
cp /tmp/bench_host_rank0.log $O/r2c5_bench_host.log 2>/dev/null
BWAGPU_HOST_DEFLATE=1 timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-aln-only --no-parity > $O/r2c5_bench_zlib.json 2> $O/r2c5_bench_zlib.err
echo "bench zlib rc=$?" >> $O/r2c5_box.log
tail -5 $O/r2c5_bgzf.log; tail -5 $O/r2c5_host.log
grep -E "host CPU|pipelined|device calls" $O/r2c5_bench_host.log | tail -4
cat $O/r2c5_box.log
