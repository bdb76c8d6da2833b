#!/bin/bash
mkdir -p gpurun_out
echo "== launch list (bench, training part)"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:^(user_pass|spmm_|adam1|user_fixup|reduce_|bias_|col_sum|relu_|gemm_f32|scale_|fill_)' -c 300 --csv --log-file gpurun_out/launches_train_v8.csv python bench.py --steps 3 --warmup 3 --topk none --no-cpu-baseline > gpurun_out/ncu_train.log 2>&1; echo "exit $?"
echo "== ncu full: spmm_seg_kernel (one whole step: 5 launches, the item pass is the long one)"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmm_seg_kernel -s 15 -c 5 -o gpurun_out/prof_item_pass_v8 -f python bench.py --steps 2 --warmup 3 --topk none --no-cpu-baseline > gpurun_out/ncu_item.log 2>&1; echo "exit $?"
