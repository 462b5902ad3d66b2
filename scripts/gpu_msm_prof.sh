#!/bin/bash
TAG=${1:-r}
OUT=gpurun_out; mkdir -p $OUT
CMD="python scripts/bench_msm.py 22"
timeout 300 $CMD > $OUT/${TAG}_msm_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_simple_shaped -s 1 -c 1 -f -o $OUT/${TAG}_prof_msm $CMD > $OUT/${TAG}_ncu_msm.log 2>&1
tail -2 $OUT/${TAG}_ncu_msm.log; cut -c1-160 $OUT/${TAG}_msm_plain.log | tail -1
