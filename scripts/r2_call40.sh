#!/bin/bash
# round 2, GPU call 40: K5 with the shorter pass-2 scans (parity + throughput), then the e2e step under lane / group settings
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "sw or global or path" > $O/r2c40_pytest.log 2>&1; echo "pytest rc=$?"
tail -2 $O/r2c40_pytest.log
timeout 600 python scripts/sw_bench.py 2>&1 | grep -E "K5 kernel" > $O/r2c40_sw_bench.log; cat $O/r2c40_sw_bench.log
rm -f /tmp/bench_host_rank0.log
timeout 900 python scripts/e2e_variants.py --steps 4 --warmup 2 base: trace:BWAGPU_TRACE=1 g3l6:BWAGPU_CALL_GROUPS=3,BWAGPU_LANES=6 g3l3:BWAGPU_CALL_GROUPS=3,BWAGPU_LANES=3 \
  g4l4:BWAGPU_CALL_GROUPS=4,BWAGPU_LANES=4 g2l2:BWAGPU_CALL_GROUPS=2,BWAGPU_LANES=2 b64k:BWAGPU_BATCH_RECORDS=65536 base2: > $O/r2c40_variants.jsonl 2> $O/r2c40_variants.err; echo "variants rc=$?"
cat $O/r2c40_variants.jsonl
cp /tmp/bench_host_rank0.log $O/r2c40_variants_host.log 2>/dev/null
