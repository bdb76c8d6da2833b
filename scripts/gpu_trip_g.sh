#!/bin/bash
mkdir -p gpurun_out
echo "== gpu tests"; timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -x > gpurun_out/test_gpu.log 2>&1; echo "exit $?"; tail -30 gpurun_out/test_gpu.log
