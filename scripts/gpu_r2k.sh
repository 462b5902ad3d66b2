#!/bin/bash
# range-gate check: (row, instance) mapping vs per-instance walk, both check modes
TAG=${1:-k}
OUT=gpurun_out; mkdir -p $OUT
for W in 0 1; do
  PG_GATES_WALK=$W timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "range_gate" > $OUT/${TAG}_pytest_walk$W.log 2>&1; tail -1 $OUT/${TAG}_pytest_walk$W.log
  for M in generic sparse; do
    PG_GATES_WALK=$W PG_CHECK_MODE=$M PG_CHECK_SHAPE=1 timeout 600 python scripts/bench_range_gate.py > $OUT/${TAG}_range_gate_walk${W}_$M.jsonl 2>> $OUT/${TAG}_range_gate.err
    echo "walk=$W mode=$M"; cut -c1-250 $OUT/${TAG}_range_gate_walk${W}_$M.jsonl
  done
done
tail -3 $OUT/${TAG}_range_gate.err
