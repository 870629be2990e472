#!/bin/bash
# round 2, GPU call 34: K5 pass 1 with the DPX cell -- the SW / path parity tests, then the K5 throughput table
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "sw or global or path" > $O/r2c34_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 $O/r2c34_pytest.log
timeout 600 python scripts/sw_bench.py > $O/r2c34_sw_bench.log 2>&1; echo "sw_bench rc=$?"
cat $O/r2c34_sw_bench.log
