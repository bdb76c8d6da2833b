#!/bin/bash
mkdir -p gpurun_out
echo "== score tests"; timeout 900 python -m pytest tests/test_gpu_score.py -q --timeout 300 -x > gpurun_out/r2_test_score.log 2>&1; echo "exit $?"; tail -4 gpurun_out/r2_test_score.log | cut -c1-300
bash scripts/r2_trip8.sh
echo "== full topk"; timeout 600 python bench.py --topk-only --topk-steps 3 2>/dev/null | python -c "import json,sys; t=json.load(sys.stdin); print('ms', round(t['ms_per_step'],2), 'frac', t['roofline']['frac'], 'parity', t['parity_check']['ok'], 'recall ms', t['recall_path'].get('ms'))"
