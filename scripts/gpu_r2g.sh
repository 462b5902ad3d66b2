#!/bin/bash
# 1 GPU: commitment-path tests (windowed fixed base), MSM bench (srs_powers rate), C5 sharded at N=1, generic ncu --set full capture + launch list
TAG=${1:-g}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "g1 or msm or srs or commit or lagrange" > $OUT/${TAG}_pytest_g1.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest_g1.log; tail -3 $OUT/${TAG}_pytest_g1.log
timeout 900 python scripts/bench_msm.py 16 20 22 > $OUT/${TAG}_msm.jsonl 2> $OUT/${TAG}_msm.err; cut -c1-300 $OUT/${TAG}_msm.jsonl; tail -2 $OUT/${TAG}_msm.err
bash scripts/gpu_multi.sh ${TAG} 1 2>&1 | grep -v "^\*\*\*\|OMP_NUM" | cut -c1-300
FULL="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
timeout 300 $FULL > $OUT/${TAG}_plain_full.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_check -s 7 -c 1 -f -o $OUT/${TAG}_prof_check_generic $FULL > $OUT/${TAG}_ncu_full.log 2>&1
tail -2 $OUT/${TAG}_ncu_full.log
timeout 300 $FULL > $OUT/${TAG}_plain_full2.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/${TAG}_launches_generic_2p24.csv $FULL > $OUT/${TAG}_ncu2.log 2>&1
