#!/bin/bash
mkdir -p gpurun_out
echo "== all gpu tests"; timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r2_test_gpu3.log 2>&1; echo "exit $?"; tail -25 gpurun_out/r2_test_gpu3.log | cut -c1-300
echo "== topk only"; timeout 600 python bench.py --topk-only --topk-steps 3 > gpurun_out/r2_topk_only.json 2> gpurun_out/r2_topk_only.err; echo "exit $?"; tail -3 gpurun_out/r2_topk_only.err
python - <<'PY'
import json
t=json.load(open('gpurun_out/r2_topk_only.json'))
print('topk ms', t.get('ms_per_step'), 'frac', t.get('roofline',{}).get('frac'), t.get('parity_check'), 'recall', t.get('recall_path'))
PY
