#!/bin/bash
mkdir -p gpurun_out
echo "== score tests"; timeout 900 python -m pytest tests/test_gpu_score.py -q --timeout 300 -x > gpurun_out/r2_test_score.log 2>&1; echo "exit $?"; tail -4 gpurun_out/r2_test_score.log | cut -c1-300
cp teamoflow_b200/csrc/libtmf.so /tmp/libtmf_prod.so
for v in ${VARIANTS:-m0 m4 m8 m16}; do
cp variants/libtmf_$v.so teamoflow_b200/csrc/libtmf.so
for d in 0 2; do
  echo -n "== $v CG2=1 DEBUG=$d: "
  TMF_TOPK_CG2=1 TMF_TOPK_DEBUG=$d timeout 300 python bench.py --topk-only --topk 151552x1000000x128x100 --topk-steps 3 2>/dev/null | python -c "import json,sys; t=json.load(sys.stdin); print('ms', round(t['ms_per_step'],2), 'parity', t['parity_check']['ok'])"
done
echo "== PROF $v"
TMF_TOPK_CG2=1 TMF_TOPK_PROF=1 timeout 300 python bench.py --topk-only --no-parity --topk 151552x1000000x128x100 --topk-steps 1 2>&1 >/dev/null | grep "tmf prof" | tail -3 | head -2 | cut -c1-420
done
cp /tmp/libtmf_prod.so teamoflow_b200/csrc/libtmf.so
echo "== full topk"; timeout 600 python bench.py --topk-only --topk-steps 3 2>/dev/null | python -c "import json,sys; t=json.load(sys.stdin); print('ms', round(t['ms_per_step'],2), 'frac', t['roofline']['frac'], 'parity', t['parity_check']['ok'], 'recall ms', t['recall_path'].get('ms'))"
