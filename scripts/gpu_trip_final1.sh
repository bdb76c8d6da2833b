#!/bin/bash
# 1-GPU visit: parity suite, default bench, reference arm, ncu launch lists and one --set full capture per top kernel
mkdir -p gpurun_out
echo "== gpu tests"; timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/test_gpu.log 2>&1; echo "exit $?"; tail -4 gpurun_out/test_gpu.log
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== full default"; timeout 1200 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "exit $?"; cat gpurun_out/bench_full.json | head -c 3000; echo
echo "== launch list (bench, training part)"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_train_v7.csv python bench.py --steps 3 --warmup 3 --topk none --no-cpu-baseline > gpurun_out/ncu_train.log 2>&1; echo "exit $?"
echo "== launch list (top-k)"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/launches_topk_v7.csv python scripts/probe_one.py 1 > gpurun_out/ncu_topk.log 2>&1; echo "exit $?"
echo "== ncu full: score_topk_kernel"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:score_topk_kernel -s 2 -c 1 -o gpurun_out/prof_topk_v7 -f python scripts/topk_small.py > gpurun_out/ncu_topk_full.log 2>&1; echo "exit $?"
echo "== ncu full: rerank_kernel"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:rerank_kernel -s 2 -c 1 -o gpurun_out/prof_rerank_v7 -f python scripts/topk_small.py > gpurun_out/ncu_rerank_full.log 2>&1; echo "exit $?"
ls -la gpurun_out/*.ncu-rep
