#!/bin/bash
# ncu --set full of the two secondary kernels changed in this series: k_check_gates (range-gate rows, generic mode) and k_ntt_pass
TAG=${1:-r05j}
OUT=gpurun_out; mkdir -p $OUT
CMD1="python scripts/bench_range_gate.py 22 64"
timeout 300 $CMD1 > $OUT/${TAG}_range_gate_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_check_gates -s 3 -c 1 -f -o $OUT/${TAG}_prof_gates $CMD1 > $OUT/${TAG}_ncu_gates.log 2>&1
tail -1 $OUT/${TAG}_ncu_gates.log
ncu -i $OUT/${TAG}_prof_gates.ncu-rep --page details > $OUT/${TAG}_k_check_gates_ncu_details.txt 2>&1
CMD2="python scripts/bench_ntt.py 24"
timeout 300 $CMD2 > $OUT/${TAG}_ntt_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ntt_pass -s 8 -c 3 -f -o $OUT/${TAG}_prof_ntt $CMD2 > $OUT/${TAG}_ncu_ntt.log 2>&1
tail -1 $OUT/${TAG}_ncu_ntt.log
ncu -i $OUT/${TAG}_prof_ntt.ncu-rep --page details > $OUT/${TAG}_k_ntt_pass_ncu_details.txt 2>&1
rm -f $OUT/${TAG}_prof_gates.ncu-rep $OUT/${TAG}_prof_ntt.ncu-rep
grep -n "Duration\|SM Busy\|Registers Per" $OUT/${TAG}_k_check_gates_ncu_details.txt $OUT/${TAG}_k_ntt_pass_ncu_details.txt | head -20
