#!/bin/bash
mkdir -p gpurun_out
echo "== train tests"; timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q --timeout 300 > gpurun_out/test_train.log 2>&1; echo "exit $?"; tail -5 gpurun_out/test_train.log
echo "== bench train"; timeout 900 python bench.py --topk none --no-cpu-baseline > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; echo "exit $?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_train.json'))
print('train ms/step', d['ms_per_step'], 'value', d['value'], 'step_roofline', d['step_roofline']['frac'])
print('phases', d['phases_ms'])
print('e2e', d['e2e']['value'])
PY
