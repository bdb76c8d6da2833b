#!/bin/bash
# N-GPU visit: weak-scaling bench at the GPU count of the box
N=$(nvidia-smi -L | wc -l)
mkdir -p gpurun_out
echo "== dist check"; timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/dist_check.py > gpurun_out/dist_check_${N}.log 2>&1; echo "exit $?"; grep -E "DIST|rror" gpurun_out/dist_check_${N}.log | tail -5
echo "== bench $N gpus"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "exit $?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_${N}gpu.json'))
print('N', d['n_gpus'], 'train ms/step', d['ms_per_step'], 'value', d['value'], 'phases', d['phases_ms'])
print('topk', {k:d['topk'].get(k) for k in ('ms_per_step','value','error')})
PY
tail -3 gpurun_out/bench_${N}gpu.err
