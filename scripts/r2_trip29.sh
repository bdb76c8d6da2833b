#!/bin/bash
cp teamoflow_b200/csrc/libtmf.so /tmp/libtmf_prod.so
for v in r1 m0 r1 m0; do
cp variants/libtmf_$v.so teamoflow_b200/csrc/libtmf.so
echo -n "== $v slice: "
TMF_TOPK_CG2=1 timeout 300 python bench.py --topk-only --topk 151552x1000000x128x100 --topk-steps 3 2>/dev/null | python -c "import json,sys; t=json.load(sys.stdin); print('ms', round(t['ms_per_step'],2), 'parity', t['parity_check']['ok'])"
echo -n "== $v full: "
timeout 600 python bench.py --topk-only --topk-steps 3 2>/dev/null | python -c "import json,sys; t=json.load(sys.stdin); print('ms', round(t['ms_per_step'],2), 'frac', t['roofline']['frac'], 'parity', t['parity_check']['ok'], 'recall ms', t['recall_path'].get('ms'))"
done
cp /tmp/libtmf_prod.so teamoflow_b200/csrc/libtmf.so
