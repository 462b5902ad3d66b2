#!/bin/bash
# ncu --set full of the structure-aware gate check (k_check<SPARSE>) at 2^22 instances
TAG=${1:-r}
OUT=gpurun_out; mkdir -p $OUT
CMD="python bench.py --log2n 22 --steps 1 --warmup 3 --no-cpu-baseline --check-mode sparse"
timeout 300 $CMD > $OUT/${TAG}_sparse_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_check -s 3 -c 1 -f -o $OUT/${TAG}_prof_sparse $CMD > $OUT/${TAG}_ncu_sparse.log 2>&1
tail -3 $OUT/${TAG}_ncu_sparse.log; cut -c1-300 $OUT/${TAG}_sparse_plain.log | tail -2
