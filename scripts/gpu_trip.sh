#!/bin/bash
mkdir -p gpurun_out
echo "== tests"; timeout 1200 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/test_gpu.log 2>&1; echo "exit $?"; tail -8 gpurun_out/test_gpu.log
echo "== memcheck (small cases)"; timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_train.py tests/test_gpu_score.py -m gpu -q --timeout 600 -k "golden or edge or degenerate or grid_golden or metrics_golden or spmm or massive" > gpurun_out/memcheck.log 2>&1; echo "exit $?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|error" gpurun_out/memcheck.log | tail -8
