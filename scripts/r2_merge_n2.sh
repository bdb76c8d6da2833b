#!/bin/bash
mkdir -p gpurun_out
echo "== 1-GPU merge tests"; timeout 600 python -m pytest tests/test_gpu_score.py -q --timeout 300 -k "merge or slab or peer" 2>&1 | tail -2
echo "== dist_check N=2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dist_check.py > gpurun_out/r2_dist_check_2.log 2>&1; echo "exit $?"; grep -E "DIST CHECK|MISMATCH|ERROR" gpurun_out/r2_dist_check_2.log | head -5
echo "== topk N=2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --topk-only --topk-steps 3 2>/dev/null | python -c "import json,sys; t=json.load(sys.stdin); print('ms', round(t['ms_per_step'],2), t['phases_ms'], 'parity', t['parity_check']['ok'])"
