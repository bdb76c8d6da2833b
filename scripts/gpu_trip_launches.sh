#!/bin/bash
# ncu launch lists (gpu__time_duration per launch) of the training step for the given workloads
mkdir -p gpurun_out
for w in $WLS; do
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_train_${w}_v8.csv python bench.py --workload $w --steps 3 --warmup 3 --topk none --no-cpu-baseline > gpurun_out/ncu_train_$w.log 2>&1; echo "$w exit $?"
python - <<PY
import csv,collections
rows=[r for r in csv.reader(open("gpurun_out/launches_train_${w}_v8.csv")) if len(r)>10]
hdr=rows[0]; ik=hdr.index("Kernel Name"); iv=hdr.index("Metric Value"); iu=hdr.index("Metric Unit")
d=collections.defaultdict(list)
for r in rows[1:]:
    v=float(r[iv].replace(",","")); u=r[iu]
    v = v/1e3 if u in ("ns","nsecond") else v*1e3 if u in ("ms","msecond") else v
    d[r[ik][:60]].append(v)
for k,v in sorted(d.items(), key=lambda kv:-sum(kv[1])): print("%-62s n=%3d mean=%9.2f us total=%10.1f us"%(k,len(v),sum(v)/len(v),sum(v)))
PY
done
