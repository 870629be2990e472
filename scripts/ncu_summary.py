#!/usr/bin/env python
"""Summarise an ncu report (gpurun_out/*.ncu-rep) into a small markdown file under profiles/.

    python scripts/ncu_summary.py gpurun_out/prof_search.ncu-rep profiles/r1_k_search_v1.md "title"

Reads the report HERE (no GPU needed): `ncu -i rep --page raw --csv` for the counters and
`--page source --print-source cuda,sass --csv` for the per-source-line share of stall samples
and executed instructions (needs -lineinfo at compile time).
"""
import csv
import io
import subprocess
import sys

RAW = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
]


def ncu(args):
    return subprocess.run(["ncu", *args], capture_output=True, text=True).stdout


def main():
    rep, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
    rows = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units, vals = rows[0], rows[1], rows[2]
    lines = [f"# {title}", "", f"Source report: `{rep}` (ncu --set full --clock-control none --import-source on).",
             "Numbers taken under the profiler are for analysis only; bench values come from bench.py.", "",
             f"Kernel: `{vals[hdr.index('Kernel Name')]}`", "", "| metric | value | unit |", "|---|---|---|"]
    for m in RAW:
        if m in hdr:
            i = hdr.index(m)
            lines.append(f"| {m} | {vals[i]} | {units[i]} |")
    rows = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"]))))
    cur, h, agg = None, None, []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif r[0] == "Line No":
            h = r
        elif h and len(r) > 8 and r[2] == "-" and r[0].isdigit():
            d = dict(zip(h[4:], r[4:]))
            try:
                agg.append((cur, int(r[0]), r[1].strip()[:90], int(d["# Samples"]), int(d["Instructions Executed"]),
                            int(d["Thread Instructions Executed"])))
            except (KeyError, ValueError):
                pass
    ts = sum(a[3] for a in agg) or 1
    ti = sum(a[4] for a in agg) or 1
    tt = sum(a[5] for a in agg)
    lines += ["", f"Source-level totals: {ts} stall samples, {ti} warp instructions, "
              f"{tt / ti:.2f} active threads per instruction.", "",
              "## Top source lines by stall samples", "", "| file:line | samples % | warp-inst % | avg lanes | source |", "|---|---|---|---|---|"]
    for a in sorted(agg, key=lambda x: -x[3])[:25]:
        lines.append(f"| {a[0]}:{a[1]} | {100 * a[3] / ts:.1f} | {100 * a[4] / ti:.1f} | {a[5] / max(a[4], 1):.1f} | `{a[2]}` |")
    lines += ["", "## Top source lines by executed warp instructions", "", "| file:line | samples % | warp-inst % | avg lanes | source |", "|---|---|---|---|---|"]
    for a in sorted(agg, key=lambda x: -x[4])[:20]:
        lines.append(f"| {a[0]}:{a[1]} | {100 * a[3] / ts:.1f} | {100 * a[4] / ti:.1f} | {a[5] / max(a[4], 1):.1f} | `{a[2]}` |")
    open(out, "w").write("\n".join(lines) + "\n")
    print("wrote", out)


if __name__ == "__main__":
    main()
