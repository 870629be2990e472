#!/bin/bash
# round 2, GPU call 47: k_search with the context sector in shared memory (default build) against plain loads; the search parity tests
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 bash scripts/ab.sh base noctx base noctx > gpurun_out/r2c47_ab.log 2>&1
cat gpurun_out/r2c47_ab.log
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "aln or golden or chunking or stats or resident or thread_pass or exact_reads or buckets" > gpurun_out/r2c47_pytest.log 2>&1; echo "pytest rc=$?"
tail -n 3 gpurun_out/r2c47_pytest.log
