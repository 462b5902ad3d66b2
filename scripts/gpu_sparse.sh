#!/bin/bash
# structure-aware kernel after a change: parity tests that hit it, bench lines for the launch shapes, one ncu --set full capture
TAG=${1:-s}
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -k "fault or poke or sparse or check_modes or range_check" > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest.log; tail -4 $OUT/${TAG}_pytest.log
for SH in 0 2 4 3 1; do
timeout 600 python bench.py --check-mode sparse --no-cpu-baseline --check-shape $SH --steps 3 > $OUT/${TAG}_bench_sparse_shape$SH.json 2>> $OUT/${TAG}_bench.err; python - <<PY
import json
d=json.load(open("$OUT/${TAG}_bench_sparse_shape$SH.json"))
print("shape $SH", d["ms_per_step"], d["kernel_ms"], d["roofline"].get("frac"))
PY
done
FULL="python bench.py --check-mode sparse --steps 1 --warmup 3 --no-cpu-baseline"
timeout 300 $FULL > $OUT/${TAG}_plain_full.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_check_prog -s 3 -c 1 -f -o $OUT/${TAG}_prof_check_prog $FULL > $OUT/${TAG}_ncu_full.log 2>&1
tail -2 $OUT/${TAG}_ncu_full.log
