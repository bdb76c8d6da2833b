#!/bin/bash
mkdir -p gpurun_out
echo "== gpu tests"; timeout 1800 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r2_test_gpu.log 2>&1; echo "exit $?"; tail -4 gpurun_out/r2_test_gpu.log | cut -c1-300
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
VARIANTS="elect guard" FULL=1 bash scripts/r2_ab.sh
