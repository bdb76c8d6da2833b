#!/bin/bash
# NOTE: keep these timeouts tight -- a multi-GPU call is charged N x its wall time, and a process that hangs at exit
# (round 1: destroy_process_group behind live CUDA graphs) burns the whole budget otherwise.
# 2-GPU visit: distributed correctness check + weak-scaling bench
mkdir -p gpurun_out
nvidia-smi -L
echo "== dist check"; timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dist_check.py > gpurun_out/dist_check.log 2>&1; echo "exit $?"; grep -E "rank|DIST|Error|error" gpurun_out/dist_check.log | tail -12
echo "== bench 2 gpus"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "exit $?"; tail -c 3000 gpurun_out/bench_2gpu.json; tail -5 gpurun_out/bench_2gpu.err
