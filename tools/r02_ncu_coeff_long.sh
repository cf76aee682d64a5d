#!/bin/bash
# ncu --set full of the 168-register k_coeff build (and k_contract) on an optics_SS bin-5 batch, final revision
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_coeff|k_contract" -s 2 -c 2 -f -o gpurun_out/r02l_coeff_long_contract_ss_bin5 \
   python tools/prof_case.py ss 4 65 > gpurun_out/l_ncu.log 2>&1; grep -c "Profiling" gpurun_out/l_ncu.log
ncu -i gpurun_out/r02l_coeff_long_contract_ss_bin5.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv, sys
rows = list(csv.reader(sys.stdin))
hdr = rows[0]
want = ['Kernel Name', 'gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__issue_active.avg.pct',
        'smsp__issue_active.avg.pct', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'sm__cycles_active.avg', 'sm__cycles_elapsed.avg', 'smsp__inst_executed.sum']
idx = [hdr.index(w) for w in want if w in hdr]
for r in rows[2:]:
    print({hdr[i]: r[i] for i in idx})
"
