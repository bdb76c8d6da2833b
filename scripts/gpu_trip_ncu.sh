#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:score_topk_kernel -s 4 -c 1 -o gpurun_out/prof_topk_v7_full -f python scripts/probe_one.py 1 > gpurun_out/ncu_topk_full2.log 2>&1; echo "exit $?"
ncu -i gpurun_out/prof_topk_v7_full.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); hdr=rows[0]; units=rows[1]; vals=rows[2]
for i,h in enumerate(hdr):
    if h in ('gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','lts__t_sector_hit_rate.pct','lts__t_bytes.sum','launch__grid_size'):
        print(h,'=',vals[i],units[i])
"
