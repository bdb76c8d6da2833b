#!/bin/bash
mkdir -p gpurun_out
cp teamoflow_b200/csrc/libtmf.so /tmp/libtmf_prod.so
cp variants/libtmf_dev.so teamoflow_b200/csrc/libtmf.so
for d in 0 1 2; do
  echo "== dev build, TMF_TOPK_DEBUG=$d, 151552 x 1M"
  TMF_TOPK_DEBUG=$d timeout 300 python bench.py --topk-only --topk 151552x1000000x128x100 --topk-steps 3 --no-parity 2>/dev/null | python -c "import json,sys; t=json.load(sys.stdin); print('ms', t['ms_per_step'], 'launches', t['gpu_launches'])"
done
cp /tmp/libtmf_prod.so teamoflow_b200/csrc/libtmf.so
echo "== plain run 37888 x 1M"; timeout 300 python bench.py --topk-only --topk 37888x1000000x128x100 --topk-steps 2 > gpurun_out/r2_topk_small.json 2> gpurun_out/r2_topk_small.err && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:score_topk_kernel -s 1 -c 1 -o gpurun_out/r2_prof_topk -f python bench.py --topk-only --topk 37888x1000000x128x100 --topk-steps 1 > gpurun_out/r2_ncu_topk.log 2>&1; echo "ncu exit $?"; tail -3 gpurun_out/r2_ncu_topk.log
