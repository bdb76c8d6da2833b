#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --topk-only > gpurun_out/topk_only.json 2> gpurun_out/topk_only.err; python -c "
import json; d=json.load(open('gpurun_out/topk_only.json')); print('ms', d['ms_per_step'], 'frac', d['roofline']['frac'])"
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_topk_v7.csv python scripts/probe_one.py 1 > gpurun_out/ncu_v7.log 2>&1; echo "exit $?"
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/launches_topk_v7.csv")) if len(r) > 10]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    v = float(r[vi].replace(",", "")); u = r[ui]
    v = v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v
    agg.setdefault(r[ki][:60], []).append(v)
for k, v in agg.items():
    print(f"  {k:60s} n={len(v):3d} total {sum(v):9.3f} ms  mean {sum(v)/len(v):8.3f}  first {v[0]:.3f} last {v[-1]:.3f}")
PY
