#!/bin/bash
# One GPU-box call: parity tests, smoke, bench at the metric size, launch list + one full ncu capture of the gate-check kernel.
# Usage (from the repo root on the box): bash scripts/gpu_round.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/${TAG}_gpu.txt 2>&1
nproc >> $OUT/${TAG}_gpu.txt; grep -m1 'model name' /proc/cpuinfo >> $OUT/${TAG}_gpu.txt
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
tail -15 $OUT/${TAG}_pytest.log
timeout 300 python __graft_entry__.py smoke > $OUT/${TAG}_smoke.log 2>&1; echo "smoke exit $?" >> $OUT/${TAG}_smoke.log
tail -3 $OUT/${TAG}_smoke.log
timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit $?"
tail -c 3000 $OUT/${TAG}_bench.json; tail -5 $OUT/${TAG}_bench.err
timeout 300 python bench.py --check-mode sparse --no-cpu-baseline > $OUT/${TAG}_bench_sparse.json 2>> $OUT/${TAG}_bench.err
SMALL="python bench.py --log2n 20 --steps 2 --warmup 3 --no-cpu-baseline"
timeout 300 $SMALL > $OUT/${TAG}_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv $SMALL > $OUT/${TAG}_ncu1.log 2>&1
timeout 300 $SMALL > $OUT/${TAG}_plain2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_check -s 3 -c 1 -f -o $OUT/${TAG}_prof_check $SMALL > $OUT/${TAG}_ncu2.log 2>&1
ls -la $OUT | tail -20
