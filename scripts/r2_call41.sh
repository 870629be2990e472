#!/bin/bash
# round 2, GPU call 41: bench with fresh outputs per step, pooled record chunks, ramped inflate calls
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
rm -f /tmp/bench_host_rank0.log
timeout 1200 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-aln-only --no-parity > $O/r2c41_bench.json 2> $O/r2c41_bench.err; echo "bench rc=$?"
cp /tmp/bench_host_rank0.log $O/r2c41_bench_host.log 2>/dev/null
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c41_bench.json'))
print('e2e', d['e2e']['value'], d['e2e']['ms_each_step_rank0'], 'value', d['value'], d['kernel_ms_per_step'])
PY
grep -E "outside the passes" $O/r2c41_bench_host.log | sed -n 8,10p
