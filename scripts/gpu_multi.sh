#!/bin/bash
# multi-GPU evidence: NCCL-behind-the-C-ABI tests, headline bench weak + strong, C5 sharded + gather bandwidth.  usage: gpu_multi.sh TAG N
TAG=${1:-m}; N=${2:-2}
OUT=gpurun_out; mkdir -p $OUT
export NCCL_DEBUG=WARN
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q > $OUT/${TAG}_pytest_multi.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest_multi.log; tail -3 $OUT/${TAG}_pytest_multi.log
timeout 600 $TR bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_weak_${N}gpu.json 2> $OUT/${TAG}_bench_weak.err; echo "weak exit $?"; cut -c1-250 $OUT/${TAG}_bench_weak_${N}gpu.json; tail -2 $OUT/${TAG}_bench_weak.err
timeout 600 $TR bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline --scaling strong > $OUT/${TAG}_bench_strong_${N}gpu.json 2> $OUT/${TAG}_bench_strong.err; echo "strong exit $?"; cut -c1-250 $OUT/${TAG}_bench_strong_${N}gpu.json; tail -2 $OUT/${TAG}_bench_strong.err
timeout 600 $TR bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline --check-mode sparse > $OUT/${TAG}_bench_weak_sparse_${N}gpu.json 2> $OUT/${TAG}_bench_weak_sparse.err; echo "weak sparse exit $?"; cut -c1-250 $OUT/${TAG}_bench_weak_sparse_${N}gpu.json
timeout 900 $TR scripts/bench_sharded.py > $OUT/${TAG}_c5_sharded_${N}gpu.raw 2> $OUT/${TAG}_c5_sharded.err; echo "c5 exit $?"; grep '^{' $OUT/${TAG}_c5_sharded_${N}gpu.raw > $OUT/${TAG}_c5_sharded_${N}gpu.jsonl; cut -c1-60,230-520 $OUT/${TAG}_c5_sharded_${N}gpu.jsonl; tail -3 $OUT/${TAG}_c5_sharded.err
