#!/bin/bash
TAG=${1:-q}
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest.log; tail -6 $OUT/${TAG}_pytest.log
timeout 600 python bench.py --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit $?"; cat $OUT/${TAG}_bench.json | cut -c1-700; tail -3 $OUT/${TAG}_bench.err
timeout 900 python scripts/bench_configs.py > $OUT/${TAG}_configs.jsonl 2> $OUT/${TAG}_configs.err; echo "configs exit $?"; cat $OUT/${TAG}_configs.jsonl | cut -c1-400; tail -5 $OUT/${TAG}_configs.err
timeout 900 python scripts/bench_configs.py --sparse > $OUT/${TAG}_configs_sparse.jsonl 2>> $OUT/${TAG}_configs.err; cat $OUT/${TAG}_configs_sparse.jsonl | cut -c1-200
