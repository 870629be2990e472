#!/bin/bash
# round 2, GPU call 4: host CPU accounting of the in-process bam2bam pipeline on C4
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
{ nvidia-smi -L; nproc; lscpu | grep -E "Model name|MHz|Socket|NUMA"; } > $O/r2c4_box.log 2>&1
timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-aln-only --no-parity > $O/r2c4_bench.json 2> $O/r2c4_bench.err
echo "bench rc=$?" >> $O/r2c4_box.log
cp /tmp/bench_host_rank0.log $O/r2c4_bench_host.log 2>/dev/null
grep -E "host CPU|pipelined" $O/r2c4_bench_host.log | tail -4
cat $O/r2c4_box.log
