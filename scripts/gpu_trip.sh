#!/bin/bash
mkdir -p gpurun_out
echo "== tests"; timeout 1200 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/test_gpu.log 2>&1; echo "exit $?"; tail -4 gpurun_out/test_gpu.log
for w in c4mini c3; do
  echo "== bench $w"; timeout 900 python bench.py --workload $w --topk none --no-cpu-baseline --steps 10 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "exit $?"
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_$w.json'))
print('$w', 'ms/step', d['ms_per_step'], 'value', d['value'], 'step_roofline', d['step_roofline']['frac'], 'e2e', d['e2e']['value'], 'phases', d['phases_ms'])
PY
  tail -2 gpurun_out/bench_$w.err
done
