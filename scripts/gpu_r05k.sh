#!/bin/bash
# N GPUs: the multi-GPU tests on this tree, then the gather of witness shards through ncclAllGather (equal shards) against the
# per-rank broadcasts (PG_GATHER_BCAST=1), 2^14 and 2^16 range_check instances per rank.  usage: gpu_r05k.sh TAG N
TAG=${1:-r05k}; N=${2:-2}
OUT=gpurun_out; mkdir -p $OUT
export NCCL_DEBUG=WARN
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q > $OUT/${TAG}_pytest_multi.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest_multi.log; tail -3 $OUT/${TAG}_pytest_multi.log
for L in 14 16; do
  for B in 0 1; do
    PG_GATHER_BCAST=$B timeout 600 $TR scripts/bench_sharded.py --steps 1 --warmup 1 --gather-log2n $L > $OUT/${TAG}_gather_${N}gpu_l${L}_b${B}.raw 2> $OUT/${TAG}_gather.err; echo "gather log2n=$L bcast=$B exit $?"
    grep '^{' $OUT/${TAG}_gather_${N}gpu_l${L}_b${B}.raw | grep "gather of witness" | python -c "import sys,json; [print('  ', d['bytes_per_rank_shard'], 'B per shard,', round(d['ms'],3), 'ms,', round(d['recv_GB_per_s_per_gpu'] or 0,1), 'GB/s received per GPU') for d in map(json.loads, sys.stdin)]"
  done
done
tail -3 $OUT/${TAG}_gather.err
