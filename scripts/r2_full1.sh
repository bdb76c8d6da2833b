#!/bin/bash
# 1-GPU visit: parity suite, smoke, default bench, reference arm
mkdir -p gpurun_out
echo "== gpu tests"; timeout 1800 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r2_test_gpu.log 2>&1; echo "exit $?"; tail -4 gpurun_out/r2_test_gpu.log | cut -c1-300
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== full default"; timeout 1500 python bench.py > gpurun_out/r2_bench_full.json 2> gpurun_out/r2_bench_full.err; echo "exit $?"; head -c 6000 gpurun_out/r2_bench_full.json; echo; tail -5 gpurun_out/r2_bench_full.err
echo "== reference arm"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; echo "exit $?"; head -c 2500 gpurun_out/r2_bench_ref.json; echo
