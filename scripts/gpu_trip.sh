#!/bin/bash
mkdir -p gpurun_out
echo "== gpu tests"; timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -x > gpurun_out/test_gpu.log 2>&1; echo "exit $?"; tail -5 gpurun_out/test_gpu.log
echo "== train legacy"; TMF_USER_PASS=legacy timeout 600 python bench.py --topk none --no-cpu-baseline > gpurun_out/bench_c3_legacy.json 2> gpurun_out/bench_c3_legacy.err; cat gpurun_out/bench_c3_legacy.json | cut -c1-600
echo "== c4mini sorted"; timeout 600 python bench.py --workload c4mini --topk none --no-cpu-baseline > gpurun_out/bench_c4mini.json 2> gpurun_out/bench_c4mini.err; cat gpurun_out/bench_c4mini.json | cut -c1-600
echo "== c2 sorted"; timeout 600 python bench.py --workload c2 --topk none --no-cpu-baseline > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; cat gpurun_out/bench_c2.json | cut -c1-400
echo "== full default"; timeout 1200 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "exit $?"; cat gpurun_out/bench_full.json
