#!/bin/bash
mkdir -p gpurun_out
echo "== score tests"; timeout 900 python -m pytest tests/test_gpu_score.py -q --timeout 300 > gpurun_out/r2_test_score.log 2>&1; echo "exit $?"; tail -4 gpurun_out/r2_test_score.log | cut -c1-300
cp teamoflow_b200/csrc/libtmf.so /tmp/libtmf_prod.so
cp variants/libtmf_dev.so teamoflow_b200/csrc/libtmf.so
for cg in 1 0; do
echo "== PROF CG2=$cg"
TMF_TOPK_CG2=$cg TMF_TOPK_PROF=1 timeout 300 python bench.py --topk-only --topk 151552x1000000x128x100 --topk-steps 1 2>&1 >/dev/null | grep "tmf prof" | tail -2 | cut -c1-600
done
cp /tmp/libtmf_prod.so teamoflow_b200/csrc/libtmf.so
