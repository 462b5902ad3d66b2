#!/bin/bash
TAG=${1:-r}; shift
OUT=gpurun_out; mkdir -p $OUT
CMD="python scripts/bench_msm.py $@"
timeout 600 $CMD > $OUT/${TAG}_msm_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_msm_launches.csv $CMD > $OUT/${TAG}_msm_ncu.log 2>&1
python3 - <<PY
import csv,collections
rows=[r for r in csv.reader(open("$OUT/${TAG}_msm_launches.csv")) if len(r)>10 and r[0].isdigit()]
agg=collections.OrderedDict()
for r in rows:
    name=r[4][:70]; v=float(r[-1].replace(',',''))
    agg.setdefault(name,[0,0.0]); agg[name][0]+=1; agg[name][1]+=v
for k,(c,t) in agg.items(): print(f"{c:5d} {t/1e6:10.3f} ms  {k}")
PY
