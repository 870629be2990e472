#!/bin/bash
# round 2, GPU call 55: reads on both sides of 255 bases through the C-ABI against the live reference (longer reads skip pass 0)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "test_aln_matches_reference_live" > gpurun_out/r2c55_pytest.log 2>&1; echo "pytest rc=$?"
tail -n 4 gpurun_out/r2c55_pytest.log
