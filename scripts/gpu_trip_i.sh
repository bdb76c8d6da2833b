#!/bin/bash
mkdir -p gpurun_out
echo "== gpu tests (score)"; timeout 1500 python -m pytest tests/test_gpu_score.py tests/test_gpu_ref_golden.py -m gpu -q --timeout 600 -x > gpurun_out/test_score.log 2>&1; echo "exit $?"; tail -8 gpurun_out/test_score.log
echo "== topk only"; timeout 900 python bench.py --topk-only > gpurun_out/topk_only.json 2> gpurun_out/topk_only.err; echo "exit $?"; python -c "
import json; d=json.load(open('gpurun_out/topk_only.json')); print('ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'exact', d['spot_check_exact'], 'recall ms', d['recall_path'])"
python scripts/topk_prof.py 151552 1000000 2>&1 | tail -2
python scripts/topk_prof.py 151552 125000 2>&1 | tail -2
