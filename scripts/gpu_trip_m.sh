#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"user_pass|spmm|adam1|user_fixup" -c 300 --csv --log-file gpurun_out/launches_train_v7.csv python bench.py --steps 3 --warmup 3 --topk none --no-cpu-baseline > gpurun_out/ncu_train.log 2>&1; echo "exit $?"
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/launches_train_v7.csv")) if len(r) > 10]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    v = float(r[vi].replace(",", "")); u = r[ui]
    v = v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v
    agg.setdefault(r[ki].split("(")[0][:70], []).append(v)
tot = sum(sum(v) for v in agg.values())
for k, v in agg.items():
    print(f"| `{k}` | {len(v)} | {sum(v)/len(v):.3f} | {100*sum(v)/tot:.1f} % |")
PY
