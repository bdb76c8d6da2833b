#!/bin/bash
mkdir -p gpurun_out
echo "== gpu tests"; timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -x > gpurun_out/test_gpu.log 2>&1; echo "exit $?"; tail -3 gpurun_out/test_gpu.log
echo "== full default"; timeout 1200 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "exit $?"; cat gpurun_out/bench_full.json
