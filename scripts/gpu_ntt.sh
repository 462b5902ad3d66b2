#!/bin/bash
TAG=${1:-r}
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fft or wire_poly" > $OUT/${TAG}_pytest_ntt.log 2>&1; tail -5 $OUT/${TAG}_pytest_ntt.log
timeout 600 python scripts/bench_ntt.py > $OUT/${TAG}_ntt.jsonl 2> $OUT/${TAG}_ntt.err; cat $OUT/${TAG}_ntt.jsonl | cut -c1-260; tail -3 $OUT/${TAG}_ntt.err
