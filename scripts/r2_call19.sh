#!/bin/bash
# round 2, GPU call 19: warp-per-member device inflate -- parity tests, kernel rate, bench with it and with the zlib workers
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_bgzf.py -m gpu -x -q > $O/r2c19_bgzf.log 2>&1; echo "bgzf rc=$?" > $O/r2c19_box.log
python - > $O/r2c19_inflate_rate.json 2> $O/r2c19_inflate_rate.err <<'PY'
import importlib, json, os, sys, time
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
bwa = importlib.import_module("network-aware-bwa_b200")
from test_bgzf import bam_like, bgzf_file
api = bwa.api
api.init([0])
data = bam_like(20000) * 20
packed = bgzf_file(data, 1)
out = {}
for form in ("0", "1"):
    os.environ["BWAGPU_INFLATE_THREAD_FORM"] = form
    if form == "1":
        break
    api.bgzf_inflate(packed[: 1 << 20] and packed)
    best = None
    for _ in range(3):
        t0 = time.perf_counter(); got, ooff, ms = api.bgzf_inflate(packed); dt = time.perf_counter() - t0
        best = ms if best is None else min(best, ms)
    assert got == data
    out["warp_per_member"] = {"members": int(len(ooff) - 1), "bytes_in": len(packed), "bytes_out": len(data), "kernel_ms": best, "gb_per_s_out": len(data) / best / 1e6, "host_call_ms": dt * 1e3}
print(json.dumps(out))
api.destroy()
PY
echo "rate rc=$?" >> $O/r2c19_box.log
BWAGPU_DEVICE_INFLATE=1 timeout 900 python -m pytest tests/test_inprocess_host.py tests/test_batched_bam2bam.py -m gpu -x -q > $O/r2c19_host.log 2>&1; echo "host rc=$?" >> $O/r2c19_box.log
BWAGPU_DEVICE_INFLATE=1 timeout 1200 python bench.py --steps 5 --warmup 2 --no-cpu-baseline --no-aln-only > $O/r2c19_bench_devinf.json 2> $O/r2c19_bench_devinf.err
echo "bench devinf rc=$?" >> $O/r2c19_box.log
cp /tmp/bench_host_rank0.log $O/r2c19_bench_host.log 2>/dev/null
timeout 1200 python bench.py --steps 5 --warmup 2 --no-cpu-baseline --no-aln-only --no-parity > $O/r2c19_bench_hostinf.json 2> $O/r2c19_bench_hostinf.err
echo "bench hostinf rc=$?" >> $O/r2c19_box.log
tail -2 $O/r2c19_bgzf.log; tail -2 $O/r2c19_host.log; cat $O/r2c19_inflate_rate.json
grep -E "pipelined|host CPU" $O/r2c19_bench_host.log | head -21 | tail -3
cat $O/r2c19_box.log
