#!/bin/bash
TAG=${1:-r}
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest.log; tail -6 $OUT/${TAG}_pytest.log
timeout 600 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit $?"; cat $OUT/${TAG}_bench.json | cut -c1-400; tail -3 $OUT/${TAG}_bench.err
timeout 600 python bench.py --check-shape 1 --no-cpu-baseline > $OUT/${TAG}_bench_shape1.json 2>> $OUT/${TAG}_bench.err; cat $OUT/${TAG}_bench_shape1.json | cut -c1-300
timeout 600 python bench.py --check-mode sparse --no-cpu-baseline > $OUT/${TAG}_bench_sparse.json 2>> $OUT/${TAG}_bench.err; cat $OUT/${TAG}_bench_sparse.json | cut -c1-300
FULL="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
timeout 300 $FULL > $OUT/${TAG}_plain_full.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_check -s 7 -c 1 -f -o $OUT/${TAG}_prof_check_2p24 $FULL > $OUT/${TAG}_ncu_full.log 2>&1
tail -3 $OUT/${TAG}_ncu_full.log
SMALL="python bench.py --log2n 20 --steps 2 --warmup 3 --no-cpu-baseline"
timeout 300 $SMALL > $OUT/${TAG}_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv $SMALL > $OUT/${TAG}_ncu1.log 2>&1
ls -la $OUT | grep ${TAG}
