#!/bin/bash
# Round evidence in one call: full GPU test suite, smoke, headline bench (both arms, both check modes), the other BASELINE
# configurations, the evaluation-domain and commitment kernels, then the launch list of C4 and one ncu --set full capture of the
# batch inversion (each only after the same command has exited 0 without ncu).  What comes back is limited to 64 MiB: one report.
TAG=${1:-f}
OUT=gpurun_out; mkdir -p $OUT
timeout 1800 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest.log; tail -4 $OUT/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/${TAG}_smoke.log 2>&1; tail -1 $OUT/${TAG}_smoke.log
timeout 600 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit $?"; cut -c1-330 $OUT/${TAG}_bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_reference.json 2>> $OUT/${TAG}_bench.err; cut -c1-200 $OUT/${TAG}_bench_reference.json
timeout 600 python bench.py --check-mode sparse --no-cpu-baseline > $OUT/${TAG}_bench_sparse.json 2>> $OUT/${TAG}_bench.err; cut -c1-200 $OUT/${TAG}_bench_sparse.json
timeout 900 python scripts/bench_configs.py > $OUT/${TAG}_configs_generic.jsonl 2> $OUT/${TAG}_configs.err; cut -c1-160 $OUT/${TAG}_configs_generic.jsonl
timeout 900 python scripts/bench_configs.py --sparse > $OUT/${TAG}_configs_sparse.jsonl 2>> $OUT/${TAG}_configs.err; cut -c1-160 $OUT/${TAG}_configs_sparse.jsonl
timeout 600 python scripts/bench_range_gate.py > $OUT/${TAG}_range_gate.jsonl 2> $OUT/${TAG}_range_gate.err; cut -c1-200 $OUT/${TAG}_range_gate.jsonl
timeout 600 python scripts/bench_ntt.py > $OUT/${TAG}_ntt.jsonl 2> $OUT/${TAG}_ntt.err; cut -c1-160 $OUT/${TAG}_ntt.jsonl
timeout 900 python scripts/bench_msm.py 16 18 20 22 > $OUT/${TAG}_msm.jsonl 2> $OUT/${TAG}_msm.err; cut -c1-200 $OUT/${TAG}_msm.jsonl
CMD="python scripts/prof_c4.py 24"
timeout 300 $CMD > $OUT/${TAG}_c4_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/${TAG}_launches_c4.csv $CMD > $OUT/${TAG}_ncu_c4.log 2>&1
timeout 300 $CMD > $OUT/${TAG}_c4_plain2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_batch_inv -s 2 -c 1 -f -o $OUT/${TAG}_prof_batch_inv $CMD > $OUT/${TAG}_ncu_c4_full.log 2>&1
tail -1 $OUT/${TAG}_ncu_c4_full.log
for f in bench configs ntt msm; do tail -n 3 $OUT/${TAG}_$f.err | cut -c1-200; done
