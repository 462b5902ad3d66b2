#!/bin/bash
# ncu --set full capture of the generic gate-check kernel of this build + the launch list of the same command (B200_PROFILING.md recipe)
TAG=${1:-r05h}
OUT=gpurun_out; mkdir -p $OUT
FULL="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
timeout 300 $FULL > $OUT/${TAG}_plain_full.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_check -s 7 -c 1 -f -o $OUT/${TAG}_prof_check_generic $FULL > $OUT/${TAG}_ncu_full.log 2>&1
tail -2 $OUT/${TAG}_ncu_full.log
ncu -i $OUT/${TAG}_prof_check_generic.ncu-rep --page details > $OUT/${TAG}_k_check_generic_ncu_details.txt 2>&1
ncu -i $OUT/${TAG}_prof_check_generic.ncu-rep --page raw --csv > $OUT/${TAG}_k_check_generic_ncu_raw.csv 2>&1
timeout 300 $FULL > $OUT/${TAG}_plain_full2.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/${TAG}_launches_generic_2p24.csv $FULL > $OUT/${TAG}_ncu2.log 2>&1
grep -c k_check $OUT/${TAG}_launches_generic_2p24.csv
ls -la $OUT/${TAG}_prof_check_generic.ncu-rep
