#!/bin/bash
# A/B of libtmf variants on ONE box at the full C5 size (1M x 1M), interleaved
cp teamoflow_b200/csrc/libtmf.so /tmp/libtmf_prod.so
for rep in 1 2; do for v in $VARIANTS; do
cp variants/libtmf_$v.so teamoflow_b200/csrc/libtmf.so
echo -n "== $v full: "
timeout 300 python bench.py --topk-only --no-parity --topk-steps 2 2>/dev/null | python -c "import json,sys; t=json.load(sys.stdin); print('ms', round(t['ms_per_step'],2), t['step_ms'])"
done; done
cp /tmp/libtmf_prod.so teamoflow_b200/csrc/libtmf.so
