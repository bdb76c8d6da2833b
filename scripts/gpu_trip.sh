#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench.  Logs go to gpurun_out/.
mkdir -p gpurun_out
echo "== tests"; timeout 1200 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/test_gpu.log 2>&1; echo "exit $?"; tail -5 gpurun_out/test_gpu.log
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "exit $?"; tail -1 gpurun_out/smoke.log
echo "== bench full"; timeout 1500 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "exit $?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_full.json'))
print('train ms/step', d['ms_per_step'], 'value', d['value'], 'step_roofline', d['step_roofline']['frac'])
print('phases', d['phases_ms'])
print('e2e', d['e2e']['value'])
print('topk', {k:d['topk'].get(k) for k in ('ms_per_step','value','spot_check_exact','error')}, d['topk'].get('roofline'))
print('cpu', d.get('cpu_baseline'))
PY
tail -3 gpurun_out/bench_full.err
