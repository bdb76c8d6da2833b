#!/bin/bash
mkdir -p gpurun_out
echo "== score tests (product)"; timeout 900 python -m pytest tests/test_gpu_score.py -q --timeout 300 -x > gpurun_out/r2_test_score.log 2>&1; echo "exit $?"; tail -2 gpurun_out/r2_test_score.log | cut -c1-200
bash scripts/r2_ab.sh
