#!/bin/bash
# round 2, GPU call 54: the committed build once more -- k_search timing + live check, the parity tests, smoke()
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout 300 bash scripts/ab.sh base > $O/r2c54_ab.log 2>&1; cat $O/r2c54_ab.log
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_batched_bam2bam.py -x -q -m gpu > $O/r2c54_pytest.log 2>&1; echo "pytest rc=$?"
tail -n 2 $O/r2c54_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2c54_smoke.log 2>&1; echo "smoke rc=$?"
tail -n 1 $O/r2c54_smoke.log
