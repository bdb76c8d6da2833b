#!/bin/bash
# training-path visit: parity tests of the training kernels, then the C3 bench without the top-k / CPU legs
mkdir -p gpurun_out
echo "== gpu train tests"; timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_ref_golden.py -m gpu -q --timeout 600 -x > gpurun_out/test_gpu_train.log 2>&1; echo "exit $?"; tail -5 gpurun_out/test_gpu_train.log
show() { python -c "
import json,sys; d=json.load(open(sys.argv[1])); print(round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['phases_ms'].items()}, round(d['roofline']['frac'],4), round(d['step_roofline']['frac'],4), '%.4g'%d['e2e']['value'])" $1; }
echo "== c3"; timeout 600 python bench.py --topk none --no-cpu-baseline > gpurun_out/bench_c3_up.json 2> gpurun_out/bench_c3_up.err; echo "exit $?"; show gpurun_out/bench_c3_up.json
for v in $VARIANTS; do
echo "== c3 $v"; env $v timeout 600 python bench.py --topk none --no-cpu-baseline > gpurun_out/bench_c3_up_$v.json 2> gpurun_out/bench_c3_up_$v.err; echo "exit $?"; show gpurun_out/bench_c3_up_$v.json
done
