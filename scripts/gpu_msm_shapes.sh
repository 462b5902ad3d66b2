#!/bin/bash
TAG=${1:-r}
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "g1 or msm or commit" > $OUT/${TAG}_pytest_msm.log 2>&1; tail -3 $OUT/${TAG}_pytest_msm.log
for s in 0 1 2 3; do
  echo "shape $s"; PG_MSM_SHAPE=$s timeout 600 python scripts/bench_msm.py 18 20 22 2>> $OUT/${TAG}_msm.err | cut -c1-170
done
tail -3 $OUT/${TAG}_msm.err
