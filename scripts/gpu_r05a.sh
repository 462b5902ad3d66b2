#!/bin/bash
# Range-gate rows in the generic check, three builds of the same source (build/variants/, made by hand with nvcc):
#   base = the tree at dd25934 (four multiplications + a 4-term dot product, reduced operands, constants by fr_add chains)
#   lazy = literal constants, f - 3 and u left unreduced (the default of this tree)
#   sqr  = -DPG_RANGE_ROW_SQUARINGS=1: every multiplication of the row as a 36-product squaring (fr_sqr16 / fr_redc16)
TAG=${1:-r05a}
OUT=gpurun_out; mkdir -p $OUT
for v in lazy sqr; do   # (sqr: the squaring build, removed from the tree after this run)
  cp build/variants/libpg_b200_$v.so plonk_gadgets_b200/libpg_b200.so
  timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fr_kernels or range_gate or cuda_library" > $OUT/${TAG}_pytest_$v.log 2>&1; echo "pytest $v exit $?"; tail -2 $OUT/${TAG}_pytest_$v.log
done
for v in base lazy sqr; do
  cp build/variants/libpg_b200_$v.so plonk_gadgets_b200/libpg_b200.so
  timeout 300 python scripts/bench_range_gate.py 24 64 254 > $OUT/${TAG}_range_gate_$v.jsonl 2> $OUT/${TAG}_range_gate_$v.err; echo "bench $v exit $?"
  python - <<EOF
import json
for l in open("$OUT/${TAG}_range_gate_$v.jsonl"):
    d = json.loads(l); print("$v", d["num_bits"], "check_ms %.3f step %.3f" % (d["check_ms"], d["ms_per_step"]))
EOF
done
cp build/variants/libpg_b200_lazy.so plonk_gadgets_b200/libpg_b200.so
