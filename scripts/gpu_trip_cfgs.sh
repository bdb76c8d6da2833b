#!/bin/bash
# training bench of the other configurations (no top-k / CPU legs)
mkdir -p gpurun_out
show() { python -c "
import json,sys; d=json.load(open(sys.argv[1])); print(round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['phases_ms'].items()}, round(d['roofline']['frac'],4), round(d['step_roofline']['frac'],4), '%.4g'%d['e2e']['value'])" $1; }
for w in $WLS; do
echo "== $w"; timeout 900 python bench.py --workload $w --topk none --no-cpu-baseline > gpurun_out/bench_${w}_v8.json 2> gpurun_out/bench_${w}_v8.err; echo "exit $?"; show gpurun_out/bench_${w}_v8.json
done
