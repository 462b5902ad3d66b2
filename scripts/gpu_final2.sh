#!/bin/bash
# Round evidence in one call: smoke, headline bench (both arms, the three check modes), the BASELINE configurations in both modes,
# range gate, evaluation domain, commitments.  (The full GPU test suite: separate call, it takes two minutes.)
TAG=${1:-f}
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/${TAG}_smoke.log 2>&1; tail -1 $OUT/${TAG}_smoke.log
timeout 600 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit $?"; cut -c1-330 $OUT/${TAG}_bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_reference.json 2>> $OUT/${TAG}_bench.err; cut -c1-200 $OUT/${TAG}_bench_reference.json
timeout 600 python bench.py --check-mode sparse --no-cpu-baseline > $OUT/${TAG}_bench_sparse.json 2>> $OUT/${TAG}_bench.err; cut -c1-200 $OUT/${TAG}_bench_sparse.json
timeout 600 python bench.py --check-mode fused --no-cpu-baseline > $OUT/${TAG}_bench_fused.json 2>> $OUT/${TAG}_bench.err; cut -c1-200 $OUT/${TAG}_bench_fused.json
timeout 900 python scripts/bench_configs.py > $OUT/${TAG}_configs_generic.jsonl 2> $OUT/${TAG}_configs.err; cut -c1-160 $OUT/${TAG}_configs_generic.jsonl
timeout 900 python scripts/bench_configs.py --sparse > $OUT/${TAG}_configs_sparse.jsonl 2>> $OUT/${TAG}_configs.err; cut -c1-160 $OUT/${TAG}_configs_sparse.jsonl
timeout 600 python scripts/bench_range_gate.py > $OUT/${TAG}_range_gate.jsonl 2> $OUT/${TAG}_range_gate.err; cut -c1-200 $OUT/${TAG}_range_gate.jsonl
PG_CHECK_MODE=sparse timeout 600 python scripts/bench_range_gate.py > $OUT/${TAG}_range_gate_sparse.jsonl 2>> $OUT/${TAG}_range_gate.err; cut -c1-200 $OUT/${TAG}_range_gate_sparse.jsonl
timeout 600 python scripts/bench_ntt.py > $OUT/${TAG}_ntt.jsonl 2> $OUT/${TAG}_ntt.err; cut -c1-160 $OUT/${TAG}_ntt.jsonl
timeout 900 python scripts/bench_msm.py 16 18 20 22 > $OUT/${TAG}_msm.jsonl 2> $OUT/${TAG}_msm.err; cut -c1-200 $OUT/${TAG}_msm.jsonl
for f in bench configs ntt msm range_gate; do tail -n 2 $OUT/${TAG}_$f.err | cut -c1-200; done
