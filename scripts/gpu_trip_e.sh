#!/bin/bash
mkdir -p gpurun_out
echo "== gpu tests (train)"; timeout 1500 python -m pytest tests/test_gpu_train.py -m gpu -q --timeout 600 > gpurun_out/test_train.log 2>&1; echo "exit $?"; tail -25 gpurun_out/test_train.log
for mode in sorted direct; do
  if [ $mode = direct ]; then export TMF_WMRB_DIRECT=1; else unset TMF_WMRB_DIRECT; fi
  for wl in c3 c4mini c2; do
    timeout 600 python bench.py --workload $wl --topk none --no-cpu-baseline > gpurun_out/bench_${wl}_$mode.json 2> gpurun_out/bench_${wl}_$mode.err; echo "exit $?"
    python - <<PY
import json
d=json.load(open('gpurun_out/bench_${wl}_$mode.json'))
print('$mode', '$wl', 'ms/step', round(d['ms_per_step'],4), 'phases', {k: round(v,3) for k,v in d['phases_ms'].items()}, 'loss', d['config']['loss_after'])
PY
  done
done
