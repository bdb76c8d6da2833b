#!/bin/bash
mkdir -p gpurun_out
echo "== full default"; timeout 1200 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "exit $?"; cat gpurun_out/bench_full.json; tail -3 gpurun_out/bench_full.err
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "exit $?"; cat gpurun_out/bench_ref.json
