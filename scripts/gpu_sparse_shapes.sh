#!/bin/bash
# structure-aware gate check: launch-shape sweep at the metric size + the check-mode parity tests
TAG=${1:-r}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "check_modes or golden" > $OUT/${TAG}_pytest_modes.log 2>&1; tail -3 $OUT/${TAG}_pytest_modes.log
for s in 0 1 2 3 4; do
  timeout 600 python bench.py --no-cpu-baseline --steps 3 --check-mode sparse --check-shape $s > $OUT/${TAG}_sparse_shape$s.json 2>> $OUT/${TAG}_sparse.err
  python -c "
import json; d=json.load(open('$OUT/${TAG}_sparse_shape$s.json')); print($s, round(d['value']/1e9,3), d['kernel_ms'], round(d['e2e']['value']/1e9,3))"
done
tail -3 $OUT/${TAG}_sparse.err
