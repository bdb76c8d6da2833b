#!/bin/bash
# N-GPU record run: reference arm under torchrun (rank 0 works, others exit 0), then the full default bench
N=${1:-2}
mkdir -p gpurun_out
echo "== reference arm N=$N"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_ref_n$N.json 2> gpurun_out/r2_bench_ref_n$N.err; echo "exit $?"; head -c 600 gpurun_out/r2_bench_ref_n$N.json; echo
echo "== bench N=$N"; timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N > gpurun_out/r2_bench_full_n$N.json 2> gpurun_out/r2_bench_full_n$N.err; echo "exit $?"; head -c 400 gpurun_out/r2_bench_full_n$N.json; echo
