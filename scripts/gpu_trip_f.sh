#!/bin/bash
mkdir -p gpurun_out
echo "== ref golden gpu tests"; timeout 900 python -m pytest tests/test_gpu_ref_golden.py -m gpu -q --timeout 600 > gpurun_out/test_refgold.log 2>&1; echo "exit $?"; tail -40 gpurun_out/test_refgold.log
