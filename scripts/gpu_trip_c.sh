#!/bin/bash
mkdir -p gpurun_out
echo "== gpu tests"; timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -x > gpurun_out/test_gpu.log 2>&1; echo "exit $?"; tail -5 gpurun_out/test_gpu.log
echo "== topk only"; timeout 900 python bench.py --topk-only > gpurun_out/topk_only.json 2> gpurun_out/topk_only.err; echo "exit $?"; cat gpurun_out/topk_only.json; tail -3 gpurun_out/topk_only.err
