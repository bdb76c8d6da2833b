#!/bin/bash
# A/B of libtmf variants (variants/libtmf_<v>.so, scratch, git-ignored) on ONE box: C3 training bench per variant
mkdir -p gpurun_out
show() { python -c "
import json,sys; d=json.load(open(sys.argv[1])); print(round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['phases_ms'].items()}, round(d['roofline']['frac'],4), round(d['step_roofline']['frac'],4), '%.4g'%d['e2e']['value'])" $1; }
for v in $VARIANTS; do
cp variants/libtmf_$v.so teamoflow_b200/csrc/libtmf.so
echo "== c3 $v"; env $ENVV timeout 600 python bench.py --topk none --no-cpu-baseline $BARGS > gpurun_out/bench_ab_$v.json 2> gpurun_out/bench_ab_$v.err; echo "exit $?"; show gpurun_out/bench_ab_$v.json
done
