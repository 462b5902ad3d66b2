#!/bin/bash
# fused mode: launch list of the step and one ncu --set full capture of the dominant kernel (RangePre<1,1>)
TAG=${1:-u}
OUT=gpurun_out; mkdir -p $OUT
FULL="python bench.py --check-mode fused --steps 1 --warmup 3 --no-cpu-baseline"
timeout 300 $FULL > $OUT/${TAG}_plain_full.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/${TAG}_launches_fused_2p24.csv $FULL > $OUT/${TAG}_ncu1.log 2>&1
timeout 300 $FULL > $OUT/${TAG}_plain_full2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_simple -s 4 -c 1 -f -o $OUT/${TAG}_prof_rangepre_fused $FULL > $OUT/${TAG}_ncu_full.log 2>&1
tail -2 $OUT/${TAG}_ncu_full.log
