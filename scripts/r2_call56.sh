#!/bin/bash
# round 2, GPU call 56: the bench line of the committed tree once more (short: K = 3, no CPU legs)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 130 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r2c56_bench.json 2> gpurun_out/r2c56_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c56_bench.json'))
r=d['roofline']
print('e2e', d['e2e']['value'], 'value', d['value'], 'frac', r['frac'], 'probe frac', r.get('own_sector_frac_of_probe'), 'l2', r.get('l2_requests'))
PY
