#!/bin/bash
mkdir -p gpurun_out
echo "== gpu tests (score)"; timeout 1500 python -m pytest tests/test_gpu_score.py tests/test_gpu_ref_golden.py -m gpu -q --timeout 600 > gpurun_out/test_score.log 2>&1; echo "exit $?"; tail -15 gpurun_out/test_score.log
for f in auto bf16 f16; do
  if [ $f = auto ]; then unset TMF_TOPK_FMT; else export TMF_TOPK_FMT=$f; fi
  timeout 900 python bench.py --topk-only > gpurun_out/topk_only_$f.json 2> gpurun_out/topk_only_$f.err; python -c "
import json; d=json.load(open('gpurun_out/topk_only_$f.json')); print('$f', 'ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'exact', d['spot_check_exact'])"
done
