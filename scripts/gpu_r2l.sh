#!/bin/bash
# chunked range pipeline: tests + end-to-end lines of the three check modes
TAG=${1:-l}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "chunked or async or fused or headline" > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest.log; tail -4 $OUT/${TAG}_pytest.log
for M in generic sparse fused; do
timeout 600 python bench.py --check-mode $M --no-cpu-baseline --steps 3 > $OUT/${TAG}_bench_$M.json 2>> $OUT/${TAG}_bench.err; python - <<PY
import json
d=json.load(open("$OUT/${TAG}_bench_$M.json"))
print("$M", round(d["value"]/1e9,2), round(d["ms_per_step"],2), d["kernel_ms"], "e2e", round(d["e2e"]["value"]/1e9,2), round(d["e2e"]["ms_per_step"],2))
PY
done
tail -3 $OUT/${TAG}_bench.err
