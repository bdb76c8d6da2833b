#!/bin/bash
mkdir -p gpurun_out
echo "== new gpu tests"; timeout 900 python -m pytest tests/test_gpu_score.py -m gpu -q --timeout 600 -x -k "bounded or peer or merge" > gpurun_out/test_new.log 2>&1; echo "exit $?"; tail -15 gpurun_out/test_new.log
echo "== probe"; timeout 900 python scripts/topk_shard_probe.py > gpurun_out/probe.log 2>&1; echo "exit $?"; tail -12 gpurun_out/probe.log
