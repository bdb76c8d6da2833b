#!/bin/bash
cp teamoflow_b200/csrc/libtmf.so /tmp/libtmf_prod.so
cp variants/libtmf_dev.so teamoflow_b200/csrc/libtmf.so
TMF_TOPK_PROF=1 timeout 300 python bench.py --topk-only --topk 151552x1000000x128x100 --topk-steps 1 2>&1 >/dev/null | grep "tmf prof" | tail -2 | cut -c1-500
cp /tmp/libtmf_prod.so teamoflow_b200/csrc/libtmf.so
timeout 300 python bench.py --topk-only --topk 151552x1000000x128x100 --topk-steps 2 > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 30 --csv --log-file gpurun_out/r2_launches_topk_slice.csv python bench.py --topk-only --topk 151552x1000000x128x100 --topk-steps 1 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2_launches_topk_slice.csv')) if len(r)>5 and r[0].isdigit()]
for r in rows[:30]: print(r[4][:60], r[-1])
PY
