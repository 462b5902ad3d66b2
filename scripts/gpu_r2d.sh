#!/bin/bash
# full GPU suite + bench lines in the three check modes + C4 with the error path
TAG=${1:-d}
OUT=gpurun_out; mkdir -p $OUT
timeout 1700 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest.log; tail -5 $OUT/${TAG}_pytest.log
for M in sparse fused; do
timeout 600 python bench.py --check-mode $M --no-cpu-baseline --steps 3 > $OUT/${TAG}_bench_$M.json 2>> $OUT/${TAG}_bench.err; python - <<PY
import json
d=json.load(open("$OUT/${TAG}_bench_$M.json"))
print("$M", d["value"], d["ms_per_step"], d["kernel_ms"], d["roofline"].get("frac"), d["e2e"]["value"])
PY
done
tail -3 $OUT/${TAG}_bench.err
