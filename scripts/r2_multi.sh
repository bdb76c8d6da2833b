#!/bin/bash
# N-GPU visit (N = $1): distributed correctness check, then the default bench under torchrun
N=${1:-2}
mkdir -p gpurun_out
echo "== dist_check N=$N"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/dist_check.py > gpurun_out/r2_dist_check_$N.log 2>&1; echo "exit $?"; grep -E "DIST CHECK|MISMATCH|ERROR|FALLBACK|Traceback|Error" gpurun_out/r2_dist_check_$N.log | head -20; grep "rank 0" gpurun_out/r2_dist_check_$N.log | head -8 | cut -c1-250
echo "== bench N=$N"; timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N ${BENCH_ARGS} > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "exit $?"; tail -c 9000 gpurun_out/r2_bench_n$N.json; echo; grep -v "^\[rank [1-9]" gpurun_out/r2_bench_n$N.err | tail -12 | cut -c1-600
