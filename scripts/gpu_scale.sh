#!/bin/bash
# weak-scaling evidence: torchrun bench.py --gpus N on N GPUs of one box
N=${1:-2}; TAG=${2:-r}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_${N}gpu.json 2> $OUT/${TAG}_bench_${N}gpu.err
echo "exit $?"; cut -c1-420 $OUT/${TAG}_bench_${N}gpu.json; tail -3 $OUT/${TAG}_bench_${N}gpu.err
