#!/bin/bash
# Range-gate evidence in one call: full GPU test suite, smoke, C++ reference tests, configuration lines (both check modes).
TAG=${1:-rg}
OUT=gpurun_out; mkdir -p $OUT
timeout 1800 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest.log; tail -4 $OUT/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/${TAG}_smoke.log 2>&1; tail -1 $OUT/${TAG}_smoke.log
timeout 300 plonk_gadgets_b200/host/reference_tests.bin > $OUT/${TAG}_cpp.log 2>&1; tail -2 $OUT/${TAG}_cpp.log
timeout 900 python scripts/bench_configs.py > $OUT/${TAG}_configs_generic.jsonl 2> $OUT/${TAG}_configs.err; grep range_gate $OUT/${TAG}_configs_generic.jsonl | cut -c1-600
timeout 900 python scripts/bench_configs.py --sparse > $OUT/${TAG}_configs_sparse.jsonl 2>> $OUT/${TAG}_configs.err; grep range_gate $OUT/${TAG}_configs_sparse.jsonl | cut -c1-600
tail -3 $OUT/${TAG}_configs.err | cut -c1-300
