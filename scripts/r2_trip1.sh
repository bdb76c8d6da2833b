#!/bin/bash
# round 2, trip 1: full GPU parity suite + smoke + default bench (C3 + C4 + C5 + parity checks) + fit() set-up profile
mkdir -p gpurun_out
nvidia-smi -L; free -g | head -2
echo "== gpu tests"; timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/r2_test_gpu.log 2>&1; echo "exit $?"; tail -8 gpurun_out/r2_test_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench"; timeout 900 python bench.py > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err; echo "exit $?"; tail -5 gpurun_out/r2_bench_1gpu.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_1gpu.json'))
print('c3 ms/step', d['ms_per_step'], 'value %.3e'%d['value'], 'e2e %.3e'%d['e2e']['value'], d['e2e']['note'][-80:])
print('phases', {k: round(v,3) for k,v in d['phases_ms'].items()})
print('parity', d['parity_check'])
print('c4', {k: d['c4'].get(k) for k in ('ms_per_step','value','phases_ms','setup_s','error')})
t=d.get('topk',{}); print('topk', t.get('ms_per_step'), t.get('roofline',{}).get('frac'), t.get('parity_check'), t.get('error'))
print('wall', d.get('bench_wall_s'))
PY
echo "== fit profile"; CUDA_LAUNCH_BLOCKING=1 timeout 300 python scripts/profile_fit.py c3 > gpurun_out/r2_profile_fit.log 2>&1; head -40 gpurun_out/r2_profile_fit.log
