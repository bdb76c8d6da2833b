#!/bin/bash
mkdir -p gpurun_out
echo "== score tests"; timeout 900 python -m pytest tests/test_gpu_score.py tests/test_gpu_ref_golden.py tests/test_gpu_baseline_configs.py -q --timeout 600 -x > gpurun_out/r2_test_score.log 2>&1; echo "exit $?"; tail -15 gpurun_out/r2_test_score.log
echo "== topk only"; timeout 600 python bench.py --topk-only --topk-steps 3 > gpurun_out/r2_topk_only.json 2> gpurun_out/r2_topk_only.err; echo "exit $?"; tail -3 gpurun_out/r2_topk_only.err
python - <<'PY'
import json
t=json.load(open('gpurun_out/r2_topk_only.json'))
print('topk ms', t.get('ms_per_step'), 'frac', t.get('roofline',{}).get('frac'), t.get('parity_check'), 'recall', t.get('recall_path'))
PY
