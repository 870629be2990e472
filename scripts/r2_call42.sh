#!/bin/bash
# round 2, GPU call 42: e2e step with the pop cap (stragglers of a small launch finish on the warp pass) and other settings
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
rm -f /tmp/bench_host_rank0.log
timeout 900 python scripts/e2e_variants.py --steps 5 --warmup 2 base: cap4096:BWAGPU_POP_CAP=4096 cap8192:BWAGPU_POP_CAP=8192 cap2048:BWAGPU_POP_CAP=2048 \
  b96k:BWAGPU_BATCH_RECORDS=98304 b192k:BWAGPU_BATCH_RECORDS=196608 base2: > $O/r2c42_variants.jsonl 2> $O/r2c42_variants.err; echo "variants rc=$?"
cat $O/r2c42_variants.jsonl
