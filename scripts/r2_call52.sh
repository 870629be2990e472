#!/bin/bash
# round 2, GPU call 52: the whole GPU test suite, smoke(), and the bench line with the request-lean k_search
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > $O/r2c52_pytest.log 2>&1; echo "pytest rc=$?"
tail -n 3 $O/r2c52_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2c52_smoke.log 2>&1; echo "smoke rc=$?"
tail -n 2 $O/r2c52_smoke.log
rm -f /tmp/bench_host_rank0.log
timeout 600 python bench.py --steps 8 --warmup 3 > $O/r2c52_bench.json 2> $O/r2c52_bench.err; echo "bench rc=$?"
cp /tmp/bench_host_rank0.log $O/r2c52_bench_host.log 2>/dev/null
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c52_bench.json'))
print('e2e', d['e2e']['value'], d['e2e']['ms_each_step_rank0'])
print('value', d['value'], d['kernel_ms_per_step'])
print('roofline', {k: d['roofline'][k] for k in ('achieved','frac','traffic','kernel_ms_per_launch','l2_requests') if k in d['roofline']})
print('cpu', d['cpu_baseline']); print('parity', d['parity_sample'])
print('k4', d['other_kernels']['k4_sa'])
PY
