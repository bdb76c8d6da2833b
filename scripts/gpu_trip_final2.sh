#!/bin/bash
# 1-GPU visit: parity suite, smoke, default bench, reference arm, training launch list, ncu --set full of the two training kernels
mkdir -p gpurun_out
echo "== gpu tests"; timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/test_gpu.log 2>&1; echo "exit $?"; tail -4 gpurun_out/test_gpu.log
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== full default"; timeout 1200 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "exit $?"; cat gpurun_out/bench_full.json | head -c 3500; echo
echo "== reference arm"; timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "exit $?"; cat gpurun_out/bench_ref.json | head -c 1200; echo
echo "== launch list (bench, training part)"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:tmf -c 300 --csv --log-file gpurun_out/launches_train_v8.csv python bench.py --steps 3 --warmup 3 --topk none --no-cpu-baseline > gpurun_out/ncu_train.log 2>&1; echo "exit $?"
echo "== ncu full: user_pass_kernel"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:user_pass_kernel -s 3 -c 1 -o gpurun_out/prof_user_pass_v8 -f python bench.py --steps 2 --warmup 3 --topk none --no-cpu-baseline > gpurun_out/ncu_up.log 2>&1; echo "exit $?"
echo "== ncu full: spmm_seg_kernel (item pass)"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmm_seg_kernel -s 14 -c 1 -o gpurun_out/prof_item_pass_v8 -f python bench.py --steps 2 --warmup 3 --topk none --no-cpu-baseline > gpurun_out/ncu_item.log 2>&1; echo "exit $?"
ls -la gpurun_out/*v8.ncu-rep
