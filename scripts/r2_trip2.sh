#!/bin/bash
mkdir -p gpurun_out
echo "== e2e probe"; timeout 300 python scripts/e2e_probe.py c3 2>&1 | tail -6
echo "== bench (no c4, no topk)"; timeout 900 python bench.py --c4 off --topk none > gpurun_out/r2_bench_c3only.json 2> gpurun_out/r2_bench_c3only.err; echo "exit $?"; tail -3 gpurun_out/r2_bench_c3only.err | cut -c1-1500
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_c3only.json'))
print('c3 ms/step', d['ms_per_step'], 'e2e %.3e'%d['e2e']['value'], d['e2e']['note'][-90:])
print('parity', d['parity_check'])
print(json.dumps(d['parity_detail'])[:3000])
PY
