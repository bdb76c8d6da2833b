#!/bin/bash
# strong-scaling C4 (10M x 2M, 500M interactions) at the GPU count of the box
N=$(nvidia-smi -L | wc -l)
mkdir -p gpurun_out
free -g | head -2
if [ "$N" = "1" ]; then
  timeout 1200 python bench.py --workload c4 --steps 10 --warmup 3 --topk none --no-cpu-baseline > gpurun_out/bench_c4_1gpu.json 2> gpurun_out/bench_c4_1gpu.err; echo "exit $?"
else
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --workload c4 --steps 10 --warmup 3 --topk none --no-cpu-baseline > gpurun_out/bench_c4_${N}gpu.json 2> gpurun_out/bench_c4_${N}gpu.err; echo "exit $?"
fi
python - <<PY
import json
d=json.load(open('gpurun_out/bench_c4_${N}gpu.json'))
print('N', d['n_gpus'], 'ms/step', round(d['ms_per_step'],3), 'value', f"{d['value']:.4e}", d['scaling'], 'phases', {k: round(v,3) for k,v in d['phases_ms'].items()})
print('e2e', d['e2e']['value'], d['config'].get('grad_exchange'), 'step_roofline', d['step_roofline']['frac'])
PY
tail -3 gpurun_out/bench_c4_${N}gpu.err
