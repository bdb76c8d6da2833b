#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --topk-only --topk 37888x1000000x128x100 --topk-steps 2 > gpurun_out/r2_topk_small.json 2> gpurun_out/r2_topk_small.err && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:score_topk_kernel -s 1 -c 1 -o gpurun_out/r2_prof_topk2 -f python bench.py --topk-only --topk 37888x1000000x128x100 --topk-steps 1 > gpurun_out/r2_ncu_topk.log 2>&1; echo "ncu exit $?"; tail -2 gpurun_out/r2_ncu_topk.log
