#!/bin/bash
cp teamoflow_b200/csrc/libtmf.so /tmp/libtmf_prod.so
for v in dev; do
cp variants/libtmf_$v.so teamoflow_b200/csrc/libtmf.so
for cg in 1 0; do for d in 0 2; do
  echo -n "== $v CG2=$cg DEBUG=$d: "
  TMF_TOPK_CG2=$cg TMF_TOPK_DEBUG=$d timeout 300 python bench.py --topk-only --topk 151552x1000000x128x100 --topk-steps 3 2>/dev/null | python -c "import json,sys; t=json.load(sys.stdin); print('ms', round(t['ms_per_step'],2), 'parity', t['parity_check']['ok'])"
done; done; done
cp /tmp/libtmf_prod.so teamoflow_b200/csrc/libtmf.so
