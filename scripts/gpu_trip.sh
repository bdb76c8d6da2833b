#!/bin/bash
mkdir -p gpurun_out
echo "== tests"; timeout 1200 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/test_gpu.log 2>&1; echo "exit $?"; tail -3 gpurun_out/test_gpu.log
for w in c1 c2 c4mini; do
  echo "== bench $w"; timeout 900 python bench.py --workload $w --topk none --no-cpu-baseline --steps 10 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "exit $?"
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_$w.json'))
print('$w', 'ms/step', d['ms_per_step'], 'value', d['value'], 'step_roofline', d['step_roofline']['frac'], 'e2e', d['e2e']['value'], 'phases', d['phases_ms'])
PY
  tail -2 gpurun_out/bench_$w.err
done
CMD="python bench.py --steps 2 --warmup 3 --topk none --no-cpu-baseline"
$CMD > gpurun_out/ncu_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'user_pass_kernel|spmm_seg_kernel|spmm_fixup_kernel|adam1_kernel' -c 120 --csv --log-file gpurun_out/launches_train_v3.csv $CMD > gpurun_out/ncu_launch.log 2>&1; echo "launch list exit $?"
