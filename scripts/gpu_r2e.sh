#!/bin/bash
# full GPU suite, the three check modes, the BASELINE configurations (C4 with its error path), launch list + SASS mix
TAG=${1:-e}
OUT=gpurun_out; mkdir -p $OUT
timeout 1700 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest.log; tail -4 $OUT/${TAG}_pytest.log
for M in sparse fused; do
timeout 600 python bench.py --check-mode $M --no-cpu-baseline --steps 3 > $OUT/${TAG}_bench_$M.json 2>> $OUT/${TAG}_bench.err; python - <<PY
import json
d=json.load(open("$OUT/${TAG}_bench_$M.json"))
print("$M", d["value"], d["ms_per_step"], d["kernel_ms"], d["roofline"].get("frac"), d["e2e"]["value"])
PY
done
tail -3 $OUT/${TAG}_bench.err
timeout 900 python scripts/bench_configs.py > $OUT/${TAG}_configs_generic.jsonl 2> $OUT/${TAG}_configs.err; cut -c1-220 $OUT/${TAG}_configs_generic.jsonl; tail -3 $OUT/${TAG}_configs.err
timeout 900 python scripts/bench_configs.py --sparse > $OUT/${TAG}_configs_sparse.jsonl 2>> $OUT/${TAG}_configs.err; cut -c1-220 $OUT/${TAG}_configs_sparse.jsonl
SMALL="python bench.py --check-mode sparse --log2n 20 --steps 2 --warmup 3 --no-cpu-baseline"
timeout 300 $SMALL > $OUT/${TAG}_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_sparse_log2n20.csv $SMALL > $OUT/${TAG}_ncu1.log 2>&1
