#!/bin/bash
# round 2, GPU call 43: DRAM sectors per random 32-byte load, by load instruction and L2 fetch-granularity limit
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
P=scripts/probe/fetch_probe
for lim in none 32 128; do timeout 120 $P $lim 3.0 3; done > $O/r2c43_probe_times.log 2>&1
cat $O/r2c43_probe_times.log
M=dram__sectors_read.sum,lts__t_requests_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum,gpu__time_duration.sum
for lim in none 32; do
  timeout 300 ncu --metrics $M --clock-control none --csv --log-file $O/r2c43_ncu_$lim.csv $P $lim 3.0 1 > $O/r2c43_ncu_$lim.log 2>&1; echo "ncu $lim rc=$?"
done
python - <<'PY'
import csv
for lim in ("none", "32"):
    rows = [r for r in csv.reader(open(f"gpurun_out/r2c43_ncu_{lim}.csv")) if len(r) > 10]
    h = rows[0]; ki = h.index("Kernel Name"); mi = h.index("Metric Name"); vi = h.index("Metric Value"); ii = h.index("ID")
    d = {}
    for r in rows[1:]:
        d.setdefault((int(r[ii]), r[ki][:40]), {})[r[mi]] = float(r[vi].replace(",", ""))
    print("limit", lim)
    for (i, k), m in sorted(d.items()):
        rq = m.get("lts__t_requests_srcunit_tex_op_read.sum", 0) or 1
        print(f"  {i:3d} {k:42s} us {m.get('gpu__time_duration.sum',0)/1e3:9.1f} req {rq:12.0f} lts_sect/req {m.get('lts__t_sectors_srcunit_tex_op_read.sum',0)/rq:5.2f} dram_sect/req {m.get('dram__sectors_read.sum',0)/rq:5.2f}")
PY
