#!/bin/bash
mkdir -p gpurun_out
for mode in plain bounded; do
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_slab8_$mode.csv python scripts/probe_one.py 8 $mode > gpurun_out/ncu_slab8_$mode.log 2>&1; echo "exit $?"
done
python - <<'PY'
import csv, collections
for mode in ("plain", "bounded"):
    rows = [r for r in csv.reader(open(f"gpurun_out/launches_slab8_{mode}.csv")) if len(r) > 10]
    hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        v = float(r[vi].replace(",", "")); u = r[ui]
        v = v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v
        agg.setdefault(r[ki][:60], []).append(v)
    print(mode)
    for k, v in agg.items():
        print(f"  {k:60s} n={len(v):3d} total {sum(v):9.3f} ms  mean {sum(v)/len(v):8.3f}")
PY
