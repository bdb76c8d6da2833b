#!/bin/bash
mkdir -p gpurun_out
cp teamoflow_b200/csrc/libtmf.so /tmp/libtmf_prod.so
for v in base c0 c200 c0e500 c0q12 c0q8 c0q8pf; do
  cp variants/libtmf_$v.so teamoflow_b200/csrc/libtmf.so
  echo -n "== $v: "
  timeout 300 python bench.py --topk-only --topk 151552x1000000x128x100 --topk-steps 3 2>/dev/null | python -c "import json,sys; t=json.load(sys.stdin); print('ms', round(t['ms_per_step'],2), 'parity', t['parity_check']['ok'])"
done
cp /tmp/libtmf_prod.so teamoflow_b200/csrc/libtmf.so
