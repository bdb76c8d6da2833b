#!/bin/bash
mkdir -p gpurun_out
echo "== gpu tests"; timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -x > gpurun_out/test_gpu.log 2>&1; echo "exit $?"; tail -4 gpurun_out/test_gpu.log
echo "== probe"; timeout 900 python scripts/topk_shard_probe.py > gpurun_out/probe.log 2>&1; echo "exit $?"; python - <<'PY'
import json
for l in open('gpurun_out/probe.log'):
    if l.startswith('{'):
        d=json.loads(l); print(d['G'], 'unbounded', round(d['main_unbounded_ms'],1), [(v['n_s'], round(v['bound_ms'],1), round(v['main_bounded_ms'],1), round(v['merge_slice_ms'],1), v['consistent_with_unbounded']) for v in d['variants']])
PY
