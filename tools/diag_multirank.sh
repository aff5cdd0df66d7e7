#!/bin/bash
# First GPU call of the next session (needs `gpurun --gpus 4`): find out why the 4-GPU weak-scaling bench of round 2 produced
# no line (open item).  Each run is bounded by bench.py's own watchdog (stack dump of every thread on
# expiry) and an outer timeout; NCCL prints its warnings.   bash tools/diag_multirank.sh [N]
N=${1:-4}; OUT=gpurun_out; mkdir -p $OUT
run() {  # tag, env assignments...
  local tag=$1; shift
  ( time env NCCL_DEBUG=WARN "$@" timeout 330 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
      --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N --steps 1 --warmup 3 --no-cpu-baseline --watchdog-s 120 ) \
      > $OUT/diag_${tag}_n$N.json 2> $OUT/diag_${tag}_n$N.err
  echo "$tag rc=$? $(tail -c 300 $OUT/diag_${tag}_n$N.json | cut -c1-200)"; grep -m3 -E "watchdog|Error|error" $OUT/diag_${tag}_n$N.err
}
run default                                   # the shipped configuration (two calls per step in the warm-up steps on > 2 ranks)
run merged_aid PAID_MERGE_AID_ANY_WORLD=1     # guidance rows inside the interpolated call on every world size
run eager PAID_SHARD_GRAPHS=0                 # multi-rank forwards launched eagerly (no captured collectives)
