#!/bin/bash
# round 2, GPU call 49: k_search with the pop cache (default build) against the in-place mask rewrite; the search parity tests
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 bash scripts/ab.sh base nopc base nopc > gpurun_out/r2c49_ab.log 2>&1
cat gpurun_out/r2c49_ab.log
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "aln or golden or chunking or stats or resident or thread_pass or exact_reads or buckets" > gpurun_out/r2c49_pytest.log 2>&1; echo "pytest rc=$?"
tail -n 3 gpurun_out/r2c49_pytest.log
