#!/bin/bash
# bash tools/gpu_scale.sh TAG N [extra bench args]: one torchrun bench line at N GPUs into gpurun_out/TAG_nN.json
TAG=$1; N=$2; shift 2
OUT=gpurun_out; mkdir -p $OUT
if [ "$N" = "1" ]; then
  timeout 600 python bench.py --gpus 1 "$@" > $OUT/${TAG}_n1.json 2> $OUT/${TAG}_n1.err
else
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29400 + N)) \
    bench.py --gpus $N "$@" > $OUT/${TAG}_n${N}.json 2> $OUT/${TAG}_n${N}.err
fi
echo "rc=$?"; tail -c 1500 $OUT/${TAG}_n${N}.json | cut -c1-1500; tail -2 $OUT/${TAG}_n${N}.err
