#!/bin/bash
cp teamoflow_b200/csrc/libtmf.so /tmp/libtmf_prod.so
for v in ${VARIANTS:-m8}; do
cp variants/libtmf_$v.so teamoflow_b200/csrc/libtmf.so
echo "== PROF $v"
TMF_TOPK_CG2=1 TMF_TOPK_PROF=1 timeout 300 python bench.py --topk-only --no-parity --topk 151552x1000000x128x100 --topk-steps 1 2>&1 >/dev/null | grep "tmf prof" | tail -4 | head -3 | cut -c1-420
done
cp /tmp/libtmf_prod.so teamoflow_b200/csrc/libtmf.so
