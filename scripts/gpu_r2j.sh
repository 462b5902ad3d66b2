#!/bin/bash
# full suite + smoke on the build with the template cache and the digit test; range gate in both check modes and launch shapes; C5 at 2^30 on one GPU; configs
TAG=${1:-j}
OUT=gpurun_out; mkdir -p $OUT
timeout 1700 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest.log; tail -4 $OUT/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/${TAG}_smoke.log 2>&1; tail -2 $OUT/${TAG}_smoke.log
for SH in 0 1 3; do PG_CHECK_SHAPE=$SH timeout 600 python scripts/bench_range_gate.py 24 64 > $OUT/${TAG}_range_gate_shape$SH.jsonl 2>> $OUT/${TAG}_range_gate.err; cut -c1-330 $OUT/${TAG}_range_gate_shape$SH.jsonl; done
timeout 600 python scripts/bench_range_gate.py > $OUT/${TAG}_range_gate.jsonl 2>> $OUT/${TAG}_range_gate.err; cut -c1-260 $OUT/${TAG}_range_gate.jsonl
timeout 900 python scripts/bench_configs.py > $OUT/${TAG}_configs_generic.jsonl 2> $OUT/${TAG}_configs.err; cut -c1-200 $OUT/${TAG}_configs_generic.jsonl | head -6
timeout 900 python scripts/bench_configs.py --sparse > $OUT/${TAG}_configs_sparse.jsonl 2>> $OUT/${TAG}_configs.err; cut -c1-200 $OUT/${TAG}_configs_sparse.jsonl | head -7
bash scripts/gpu_c5_big.sh ${TAG} 1 2>&1 | grep -v "^\*\*\*\|OMP_NUM" | cut -c1-40,230-420
