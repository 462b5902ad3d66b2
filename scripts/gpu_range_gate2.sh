#!/bin/bash
# Range-gate evidence: GPU tests of the range gate, throughput per width, launch list and one ncu --set full capture of k_check_gates.
TAG=${1:-rg}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q -k "range_gate or golden or cpp or permutation" > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest.log; tail -3 $OUT/${TAG}_pytest.log
timeout 600 python scripts/bench_range_gate.py > $OUT/${TAG}_range_gate.jsonl 2> $OUT/${TAG}_range_gate.err; cut -c1-420 $OUT/${TAG}_range_gate.jsonl
CMD="python scripts/bench_range_gate.py 22 64"
timeout 300 $CMD > $OUT/${TAG}_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/${TAG}_launches.csv $CMD > $OUT/${TAG}_ncu1.log 2>&1
timeout 300 $CMD > $OUT/${TAG}_plain2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_check_gates -s 2 -c 1 -f -o $OUT/${TAG}_prof_check_gates $CMD > $OUT/${TAG}_ncu2.log 2>&1
tail -2 $OUT/${TAG}_ncu2.log; tail -3 $OUT/${TAG}_range_gate.err | cut -c1-300
