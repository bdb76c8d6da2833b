#!/bin/bash
# A/B of libtmf variants on ONE box: slice (151,552 x 1M) whole-call time, each variant twice, interleaved
cp teamoflow_b200/csrc/libtmf.so /tmp/libtmf_prod.so
for rep in 1 2; do for v in $VARIANTS; do
cp variants/libtmf_$v.so teamoflow_b200/csrc/libtmf.so
echo -n "== $v slice: "
timeout 300 python bench.py --topk-only --topk 151552x1000000x128x100 --topk-steps 3 2>/dev/null | python -c "import json,sys; t=json.load(sys.stdin); print('ms', round(t['ms_per_step'],2), 'parity', t['parity_check']['ok'])"
done; done
cp /tmp/libtmf_prod.so teamoflow_b200/csrc/libtmf.so
if [ -n "$FULL" ]; then echo -n "== prod full: "; timeout 600 python bench.py --topk-only --topk-steps 3 2>/dev/null | python -c "import json,sys; t=json.load(sys.stdin); print('ms', round(t['ms_per_step'],2), 'frac', t['roofline']['frac'], 'parity', t['parity_check']['ok'])"; fi
