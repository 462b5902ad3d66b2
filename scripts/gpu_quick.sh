#!/bin/bash
# quick GPU call: micro-benchmarks, parity tests, one bench line
TAG=${1:-q}
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python scripts/ubench.py > $OUT/${TAG}_ubench.json 2> $OUT/${TAG}_ubench.err; cat $OUT/${TAG}_ubench.json | tr -d '\n' | cut -c1-1500; echo; tail -3 $OUT/${TAG}_ubench.err
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest.log; tail -8 $OUT/${TAG}_pytest.log
timeout 600 python bench.py --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit $?"; cat $OUT/${TAG}_bench.json | cut -c1-2500; tail -3 $OUT/${TAG}_bench.err
