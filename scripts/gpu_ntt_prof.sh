#!/bin/bash
TAG=${1:-r}
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fft or wire_poly" > $OUT/${TAG}_pytest_ntt.log 2>&1; tail -3 $OUT/${TAG}_pytest_ntt.log
CMD="python scripts/bench_ntt.py 24"
timeout 300 $CMD > $OUT/${TAG}_ntt_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ntt_pass -s 4 -c 2 -f -o $OUT/${TAG}_prof_ntt $CMD > $OUT/${TAG}_ncu_ntt.log 2>&1
tail -2 $OUT/${TAG}_ncu_ntt.log; cut -c1-200 $OUT/${TAG}_ntt_plain.log | head -3
