#!/bin/bash
# 1-GPU profiling visit of round 2 (every ncu command after the same command exited 0 without ncu)
mkdir -p gpurun_out
TRAIN="python bench.py --steps 3 --warmup 3 --topk none --no-cpu-baseline --no-parity --no-e2e --c4 off"
echo "== plain runs"
timeout 600 $TRAIN > gpurun_out/r02_train_plain.json 2>/dev/null; echo "train exit $?"
timeout 600 python scripts/probe_one.py 1 > /dev/null 2>&1; echo "probe_one exit $?"
timeout 600 python scripts/topk_small.py > /dev/null 2>&1; echo "topk_small exit $?"
echo "== launch list (training)"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_train_c3.csv $TRAIN > gpurun_out/ncu_train.log 2>&1; echo "exit $?"
echo "== launch list (top-k 1M x 1M)"; timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 100 --csv --log-file gpurun_out/r02_launches_topk.csv python scripts/probe_one.py 1 > gpurun_out/ncu_topk.log 2>&1; echo "exit $?"
echo "== ncu full: score_topk_kernel"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:score_topk_kernel -s 2 -c 1 -o gpurun_out/r02_prof_topk -f python scripts/topk_small.py > gpurun_out/ncu_topk_full.log 2>&1; echo "exit $?"
echo "== ncu full: user_pass_kernel"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:user_pass_kernel -s 4 -c 1 -o gpurun_out/r02_prof_user_pass -f $TRAIN > gpurun_out/ncu_up_full.log 2>&1; echo "exit $?"
echo "== ncu full: spmm_seg_kernel (one step's launches)"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmm_seg -s 20 -c 5 -o gpurun_out/r02_prof_spmm -f $TRAIN > gpurun_out/ncu_spmm_full.log 2>&1; echo "exit $?"
ls -la gpurun_out/*.ncu-rep
