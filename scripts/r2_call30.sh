#!/bin/bash
# round 2, GPU call 30: compute-sanitizer memcheck over the new kernels' parity tests (BGZF deflate / inflate, K6 warp form)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 --log-file $O/r2c30_memcheck.log python -m pytest tests/test_bgzf.py tests/test_gpu_parity.py -m gpu -x -q -k "device or global_align or mate_sw_path" > $O/r2c30_pytest.log 2>&1
echo "memcheck rc=$?"
tail -3 $O/r2c30_pytest.log; tail -5 $O/r2c30_memcheck.log
