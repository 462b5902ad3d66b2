#!/bin/bash
# k_batch_inv launch shapes (blocks/SM) on C4 and on the headline's witness generation; parity first.
TAG=${1:-invs}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q -k "golden or non_zero or maybe or range_check or inv" > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest.log; tail -3 $OUT/${TAG}_pytest.log
for s in 2 3 4; do
  echo "shape $s"
  PG_INV_SHAPE=$s timeout 120 python scripts/prof_c4.py 24 2>&1 | tail -1
  PG_INV_SHAPE=$s timeout 120 python scripts/prof_c4.py 20 2>&1 | tail -1
  PG_INV_SHAPE=$s timeout 300 python bench.py --no-cpu-baseline --steps 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['kernel_ms'])"
done 2>&1 | tee $OUT/${TAG}_shapes.log
