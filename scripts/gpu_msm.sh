#!/bin/bash
TAG=${1:-r}
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "g1 or msm or commit" > $OUT/${TAG}_pytest_msm.log 2>&1; tail -5 $OUT/${TAG}_pytest_msm.log
timeout 900 python scripts/bench_msm.py > $OUT/${TAG}_msm.jsonl 2> $OUT/${TAG}_msm.err; cat $OUT/${TAG}_msm.jsonl | cut -c1-300; tail -3 $OUT/${TAG}_msm.err
