#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench, ncu.  Logs go to gpurun_out/.
mkdir -p gpurun_out
echo "== tests"; timeout 1200 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/test_gpu.log 2>&1; echo "exit $?"; tail -15 gpurun_out/test_gpu.log
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "exit $?"; tail -1 gpurun_out/smoke.log
echo "== bench full"; timeout 1500 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "exit $?"; cat gpurun_out/bench_full.json; tail -5 gpurun_out/bench_full.err
echo "== ncu"
CMD="python bench.py --steps 2 --warmup 3 --topk none --no-cpu-baseline"
$CMD > gpurun_out/ncu_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'user_pass_kernel|spmm_seg_kernel|spmm_fixup_kernel|adam1_kernel' -c 120 --csv --log-file gpurun_out/launches_train.csv $CMD > gpurun_out/ncu_launch.log 2>&1; echo "launch list exit $?"
ncu --set full --clock-control none --import-source on -k regex:'user_pass_kernel' -s 3 -c 1 -o gpurun_out/prof_user_pass $CMD > gpurun_out/ncu_user.log 2>&1; echo "user_pass capture exit $?"
ncu --set full --clock-control none --import-source on -k regex:'spmm_seg_kernel' -s 12 -c 1 -o gpurun_out/prof_item_pass $CMD > gpurun_out/ncu_item.log 2>&1; echo "spmm capture exit $?"
ls -la gpurun_out
