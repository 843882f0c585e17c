#!/bin/bash
# MFCC kernel: parity tests, stand-alone timing, one ncu --set full capture (F = 100 case).
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k mfcc > gpurun_out/pytest_mfcc.log 2>&1; echo "== pytest exit $?"; tail -n 15 gpurun_out/pytest_mfcc.log
bash scripts/gpu_ncu_mfcc.sh
python - <<'E'
import csv
rows = list(csv.reader(open("gpurun_out/prof_mfcc_raw.csv")))
hdr, vals = rows[0], rows[2]
for k in ("gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
          "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
          "launch__occupancy_limit_registers", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
          "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
          "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
          "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"):
    if k in hdr: print(k, vals[hdr.index(k)])
E
