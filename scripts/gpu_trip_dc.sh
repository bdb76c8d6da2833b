#!/bin/bash
N=$(nvidia-smi -L | wc -l)
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/dist_check.py > gpurun_out/dist_check_${N}.log 2>&1; echo "exit $?"; grep -E "ragged|DIST|rror|MISMATCH" gpurun_out/dist_check_${N}.log | tail -30
