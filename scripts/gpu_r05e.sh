#!/bin/bash
# Generic gate check, variants of two details (build/variants/, made by hand with nvcc -D...):
#   v0 = the tree as committed;  v4 = -DPG_ZERO_WIRE_LOAD=1 (the zero variable's wire is loaded like any other scalar instead of eight
#   register initialisations);  v8 = -DPG_KQ_TABLE=1 (the "0 mod q" test compares with a shared-memory table of k*q instead of
#   multiplying k*q out);  v48 = both
TAG=${1:-r05e}
OUT=gpurun_out; mkdir -p $OUT
for v in v0 v4 v8 v48 v0; do
  cp build/variants/libpg_b200_$v.so plonk_gadgets_b200/libpg_b200.so
  timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_$v.json 2> $OUT/${TAG}_bench_$v.err; echo "bench $v exit $?"
  python - <<PY
import json
d = json.loads(open("$OUT/${TAG}_bench_$v.json").read().strip().splitlines()[-1])
print("$v", "ms_per_step %.2f check %.2f witness %.2f frac %.4f" % (d["ms_per_step"], d["kernel_ms"]["check"], d["kernel_ms"]["witness"], d["roofline"]["frac"]))
PY
done
cp build/variants/libpg_b200_v48.so plonk_gadgets_b200/libpg_b200.so
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "golden or fault or headline or range_check or full_size or poked" > $OUT/${TAG}_pytest_v48.log 2>&1; echo "pytest v48 exit $?"; tail -2 $OUT/${TAG}_pytest_v48.log
cp build/variants/libpg_b200_v0.so plonk_gadgets_b200/libpg_b200.so
