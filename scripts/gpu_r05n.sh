#!/bin/bash
# Generic gate check: L1 prefetch of the next row's wires -- both kinds of wire (p0, the tree), scalar wires only (p1), none (p2) -- and the
# launch shapes 0 / 2 (20 / 24 warps per SM) on the leaner row loop of run r05e.  Builds by hand: nvcc -DPG_PREFETCH_MODE=1|2.
TAG=${1:-r05n}
OUT=gpurun_out; mkdir -p $OUT
run() {  # name lib shape
  cp build/variants/libpg_b200_$2.so plonk_gadgets_b200/libpg_b200.so
  timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --check-shape $3 > $OUT/${TAG}_bench_$1.json 2> $OUT/${TAG}_bench_$1.err; echo "bench $1 exit $?"
  python - <<PY
import json
d = json.loads(open("$OUT/${TAG}_bench_$1.json").read().strip().splitlines()[-1])
print("$1", "ms_per_step %.2f check %.2f witness %.2f" % (d["ms_per_step"], d["kernel_ms"]["check"], d["kernel_ms"]["witness"]))
PY
}
run p0 p0 0; run p1 p1 0; run p2 p2 0; run p0_shape2 p0 2; run p1_shape2 p1 2; run p0_again p0 0
cp build/variants/libpg_b200_p0.so plonk_gadgets_b200/libpg_b200.so
