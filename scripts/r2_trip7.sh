#!/bin/bash
mkdir -p gpurun_out
echo "== score tests (CG2 default)"; timeout 900 python -m pytest tests/test_gpu_score.py tests/test_gpu_ref_golden.py -q --timeout 300 -x > gpurun_out/r2_test_score.log 2>&1; echo "exit $?"; tail -6 gpurun_out/r2_test_score.log | cut -c1-300
cp teamoflow_b200/csrc/libtmf.so /tmp/libtmf_prod.so
cp variants/libtmf_dev.so teamoflow_b200/csrc/libtmf.so
for cg in 1 0; do for d in 0 2; do
  echo -n "== dev CG2=$cg DEBUG=$d 151552 x 1M: "
  TMF_TOPK_CG2=$cg TMF_TOPK_DEBUG=$d timeout 300 python bench.py --topk-only --topk 151552x1000000x128x100 --topk-steps 3 2>/dev/null | python -c "import json,sys; t=json.load(sys.stdin); print('ms', round(t['ms_per_step'],2), 'parity', t['parity_check']['ok'])"
done; done
cp /tmp/libtmf_prod.so teamoflow_b200/csrc/libtmf.so
echo "== full topk"; timeout 600 python bench.py --topk-only --topk-steps 3 2>/dev/null | python -c "import json,sys; t=json.load(sys.stdin); print('ms', round(t['ms_per_step'],2), 'frac', t['roofline']['frac'], 'parity', t['parity_check']['ok'], 'recall ms', t['recall_path'].get('ms'))"
