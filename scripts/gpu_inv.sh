#!/bin/bash
# Batch-inversion change: full GPU parity suite, then the configurations that lean on it (C4, C5) and the headline.
TAG=${1:-inv}
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest.log; tail -4 $OUT/${TAG}_pytest.log
timeout 900 python scripts/bench_configs.py > $OUT/${TAG}_configs_generic.jsonl 2> $OUT/${TAG}_configs.err; cut -c1-330 $OUT/${TAG}_configs_generic.jsonl
timeout 900 python scripts/bench_configs.py --sparse > $OUT/${TAG}_configs_sparse.jsonl 2>> $OUT/${TAG}_configs.err; cut -c1-330 $OUT/${TAG}_configs_sparse.jsonl
timeout 600 python bench.py --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit $?"; cut -c1-330 $OUT/${TAG}_bench.json
tail -3 $OUT/${TAG}_configs.err $OUT/${TAG}_bench.err | cut -c1-200
