#!/bin/bash
cp teamoflow_b200/csrc/libtmf.so /tmp/libtmf_prod.so
cp variants/libtmf_dev.so teamoflow_b200/csrc/libtmf.so
for cg in 1; do for d in 0 1 5; do
  echo -n "== dev CG2=$cg DEBUG=$d: "
  TMF_TOPK_CG2=$cg TMF_TOPK_DEBUG=$d timeout 300 python bench.py --topk-only --no-parity --topk 151552x1000000x128x100 --topk-steps 3 2>/dev/null | python -c "import json,sys; t=json.load(sys.stdin); print('ms', round(t['ms_per_step'],2))"
done; done
cp /tmp/libtmf_prod.so teamoflow_b200/csrc/libtmf.so
