#!/bin/bash
# round 2, first call: GPU test suite of the restored tree + headline bench lines in both check modes
TAG=${1:-r4a}
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest.log; tail -6 $OUT/${TAG}_pytest.log
timeout 600 python bench.py --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit $?"; cut -c1-400 $OUT/${TAG}_bench.json; tail -3 $OUT/${TAG}_bench.err
timeout 600 python bench.py --check-mode sparse --no-cpu-baseline > $OUT/${TAG}_bench_sparse.json 2>> $OUT/${TAG}_bench.err; cut -c1-300 $OUT/${TAG}_bench_sparse.json
