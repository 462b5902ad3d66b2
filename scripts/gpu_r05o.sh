#!/bin/bash
# The tree without the L1 prefetch in the generic check (np): parity / fault-injection tests, bench; and the range-gate walk kernel
# (structure-aware mode) with and without its prefetch (npg = -DPG_GATES_NO_PREFETCH=1)
TAG=${1:-r05o}
OUT=gpurun_out; mkdir -p $OUT
cp build/variants/libpg_b200_np.so plonk_gadgets_b200/libpg_b200.so
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "golden or fault or headline or range_check or full_size or poked" > $OUT/${TAG}_pytest_np.log 2>&1; echo "pytest np exit $?"; tail -2 $OUT/${TAG}_pytest_np.log
timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_np.json 2> $OUT/${TAG}_bench_np.err; echo "bench exit $?"; cut -c1-170 $OUT/${TAG}_bench_np.json
for v in np npg; do
  cp build/variants/libpg_b200_$v.so plonk_gadgets_b200/libpg_b200.so
  PG_CHECK_MODE=sparse timeout 300 python scripts/bench_range_gate.py 24 64 254 > $OUT/${TAG}_range_gate_sparse_$v.jsonl 2> $OUT/${TAG}_rg_$v.err
  python - <<PY
import json
for l in open("$OUT/${TAG}_range_gate_sparse_$v.jsonl"):
    d = json.loads(l); print("$v sparse walk", d["num_bits"], "check_ms %.3f" % d["check_ms"])
PY
done
cp build/variants/libpg_b200_np.so plonk_gadgets_b200/libpg_b200.so
