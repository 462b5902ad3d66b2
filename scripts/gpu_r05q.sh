#!/bin/bash
# ncu --set full of the generic gate-check kernel of the final tree (no L1 prefetch)
TAG=${1:-r05q}
OUT=gpurun_out; mkdir -p $OUT
FULL="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
timeout 200 $FULL > $OUT/${TAG}_plain_full.log 2>&1 &&
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_check -s 7 -c 1 -f -o $OUT/${TAG}_prof_check_generic $FULL > $OUT/${TAG}_ncu_full.log 2>&1
tail -1 $OUT/${TAG}_ncu_full.log
ncu -i $OUT/${TAG}_prof_check_generic.ncu-rep --page details > $OUT/${TAG}_k_check_generic_ncu_details.txt 2>&1
ncu -i $OUT/${TAG}_prof_check_generic.ncu-rep --page raw --csv > $OUT/${TAG}_k_check_generic_ncu_raw.csv 2>&1
rm -f $OUT/${TAG}_prof_check_generic.ncu-rep
grep -n "Duration\|SM Busy\|Issue Slots Busy" $OUT/${TAG}_k_check_generic_ncu_details.txt | head -4
