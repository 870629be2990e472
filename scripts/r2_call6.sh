#!/bin/bash
# round 2, GPU call 6: pipeline shape sweep (batch size x step size) after the to_seq / destroy stage changes
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
: > $O/r2c6_box.log
for cfg in "500000 65536" "500000 131072" "2000000 65536" "2000000 131072" "2000000 262144"; do
  set -- $cfg
  BWAGPU_BATCH_RECORDS=$2 timeout 900 python bench.py --steps 3 --warmup 2 --pairs $1 --no-cpu-baseline --no-aln-only --no-parity > $O/r2c6_p$1_b$2.json 2> $O/r2c6_p$1_b$2.err
  echo "pairs $1 batch $2 rc=$?" >> $O/r2c6_box.log
  grep -E "pipelined" /tmp/bench_host_rank0.log | tail -2 >> $O/r2c6_box.log
  grep -E "host CPU" /tmp/bench_host_rank0.log | tail -1 >> $O/r2c6_box.log
done
cat $O/r2c6_box.log
