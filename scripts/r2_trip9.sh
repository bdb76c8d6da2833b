#!/bin/bash
cp teamoflow_b200/csrc/libtmf.so /tmp/libtmf_prod.so
cp variants/libtmf_dev.so teamoflow_b200/csrc/libtmf.so
for cg in 0 1; do for d in 2 0; do
  echo "== dev PROF CG2=$cg DEBUG=$d 37888 x 1M"
  TMF_TOPK_PROF=1 TMF_TOPK_CG2=$cg TMF_TOPK_DEBUG=$d timeout 300 python bench.py --topk-only --topk 37888x1000000x128x100 --topk-steps 2 2>&1 >/dev/null | grep "tmf prof" | tail -2 | cut -c1-400
done; done
cp /tmp/libtmf_prod.so teamoflow_b200/csrc/libtmf.so
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw --format=csv
