#!/bin/bash
# round 2, GPU call 57: bench.py with the parity leg off (the roofline leg then sets the device up itself); small workload
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 65 python bench.py --steps 1 --warmup 3 --pairs 200000 --aln-reads 500000 --no-cpu-baseline --no-parity > gpurun_out/r2c57_bench.json 2> gpurun_out/r2c57_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2c57_bench.err
python -c "
import json
d=json.load(open('gpurun_out/r2c57_bench.json')); r=d['roofline']
print('e2e', d['e2e']['value'], 'value', d['value'], 'frac', r['frac'], 'probe frac', r.get('own_sector_frac_of_probe'), 'l2', r.get('l2_requests'))"
