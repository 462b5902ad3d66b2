#!/bin/bash
TAG=${1:-n}
OUT=gpurun_out; mkdir -p $OUT
for SH in 0 2 1; do
for rep in 1 2; do
timeout 600 python bench.py --check-mode sparse --no-cpu-baseline --check-shape $SH --steps 5 > $OUT/${TAG}_bench_sparse_shape${SH}_$rep.json 2>> $OUT/${TAG}_bench.err; python - <<PY
import json
d=json.load(open("$OUT/${TAG}_bench_sparse_shape${SH}_$rep.json"))
print("shape $SH rep $rep", round(d["ms_per_step"],3), d["kernel_ms"], d["roofline"].get("frac"))
PY
done; done
