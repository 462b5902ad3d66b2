#!/bin/bash
# prefetch variants of the structure-aware row program (libraries prebuilt under build/variants): bench --check-mode sparse with each
OUT=gpurun_out; mkdir -p $OUT
for v in base a4 a6 a20 l2; do
  cp build/variants/libpg_$v.so plonk_gadgets_b200/libpg_b200.so
  echo "variant $v: $(timeout 300 python bench.py --check-mode sparse --no-cpu-baseline --steps 4 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['kernel_ms'])")"
done 2>&1 | tee $OUT/${1:-v}_sp_variants.log
