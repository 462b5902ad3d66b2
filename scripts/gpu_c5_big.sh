#!/bin/bash
# C5's mix at 2^30 rows (16 x the BASELINE size: a step that is not bounded by per-call latency), sharded by pg_shard_plan.  usage: gpu_c5_big.sh TAG N
TAG=${1:-c}; N=${2:-8}
OUT=gpurun_out; mkdir -p $OUT
export NCCL_DEBUG=WARN
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519"
timeout 900 $TR scripts/bench_sharded.py --log2rows 30 --steps 2 --warmup 1 --gather-log2n 17 > $OUT/${TAG}_c5_2p30_${N}gpu.raw 2> $OUT/${TAG}_c5_2p30.err; echo "c5 big exit $?"; grep '^{' $OUT/${TAG}_c5_2p30_${N}gpu.raw > $OUT/${TAG}_c5_2p30_${N}gpu.jsonl; cut -c1-40,230-560 $OUT/${TAG}_c5_2p30_${N}gpu.jsonl; tail -3 $OUT/${TAG}_c5_2p30.err
