#!/bin/bash
# round 2, GPU call 3: k_search with the compact context (ctx16), histograms, pop-cap sweep, small-batch behaviour
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
K="python scripts/kbench.py"
C4="--genome-bp 3100000000 --read-len 100 --reads 4000000 --batches 131072,524288"
C2="--genome-bp 100000000 --read-len 76 --reads 10000000"
python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $O/r2c3_parity.log 2>&1; echo "parity rc=$?" > $O/r2c3_box.log
BWAGPU_PRINT_HIST=1 $K $C4 --stats --check 20000 --tag c4_default > $O/r2c3_c4.json 2> $O/r2c3_c4.err; echo "c4 rc=$?" >> $O/r2c3_box.log
for cap in 16384 8192 4096 2048; do
  BWAGPU_POP_CAP=$cap $K $C4 --check 20000 --tag c4_cap$cap > $O/r2c3_c4_cap$cap.json 2> $O/r2c3_c4_cap$cap.err; echo "cap$cap rc=$?" >> $O/r2c3_box.log
done
BWAGPU_PRINT_HIST=1 $K $C2 --stats --check 20000 --tag c2_default > $O/r2c3_c2.json 2> $O/r2c3_c2.err; echo "c2 rc=$?" >> $O/r2c3_box.log
BWAGPU_POP_CAP=4096 $K $C2 --check 20000 --tag c2_cap4096 > $O/r2c3_c2_cap4096.json 2> $O/r2c3_c2_cap4096.err; echo "c2cap rc=$?" >> $O/r2c3_box.log
cat $O/r2c3_box.log; tail -3 $O/r2c3_parity.log
