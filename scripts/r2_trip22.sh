#!/bin/bash
cp teamoflow_b200/csrc/libtmf.so /tmp/libtmf_prod.so
cp variants/libtmf_dev.so teamoflow_b200/csrc/libtmf.so
for d in 0; do
echo "== PROF CG2=1 DEBUG=$d"
TMF_TOPK_CG2=1 TMF_TOPK_DEBUG=$d TMF_TOPK_PROF=1 timeout 300 python bench.py --topk-only --no-parity --topk 151552x1000000x128x100 --topk-steps 1 2>&1 >/dev/null | grep "tmf prof" | tail -3 | cut -c1-500
done
cp /tmp/libtmf_prod.so teamoflow_b200/csrc/libtmf.so
