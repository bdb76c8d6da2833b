#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:user_pass_kernel -s 3 -c 1 -o gpurun_out/prof_user_pass_v8 -f python bench.py --steps 2 --warmup 3 --topk none --no-cpu-baseline > gpurun_out/ncu_up.log 2>&1; echo "exit $?"
ls -la gpurun_out/*.ncu-rep
