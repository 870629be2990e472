#!/bin/bash
# round 2, GPU call 48: k_search with the top record in registers (BWAGPU_TOP_REG) on the request-lean kernel; ncu counters of the default build
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 bash scripts/ab.sh base topreg base topreg > gpurun_out/r2c48_ab.log 2>&1
cat gpurun_out/r2c48_ab.log
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_requests_srcunit_tex_op_read.sum,lts__t_requests_srcunit_tex_op_write.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sector_hit_rate.pct,l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_miss.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_global_op_st.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread
timeout 600 ncu --metrics $M --clock-control none -k regex:k_search -c 1 --csv --log-file gpurun_out/r2c48_ncu_search.csv \
  python scripts/kbench.py --genome-bp 3100000000 --read-len 100 --reads 2000000 --reps 1 > gpurun_out/r2c48_ncu_search.log 2>&1; echo "ncu rc=$?"
cut -d, -f5,13- gpurun_out/r2c48_ncu_search.csv | tail -n 20
