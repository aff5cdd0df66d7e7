#!/bin/bash
# Four-GPU pass on the final tree: multi-rank NCCL parity (world 2 and 4), weak-scaling bench line at 4 GPUs.
OUT=gpurun_out; T=${1:-r2l}
mkdir -p $OUT
( time timeout 600 python -m pytest tests/test_multirank_gpu.py -m gpu -q ) > $OUT/${T}_pytest_multirank.log 2>&1; echo "pytest multirank rc=$?" | tee -a $OUT/${T}_pytest_multirank.log; tail -4 $OUT/${T}_pytest_multirank.log
bash tools/gpu_scale.sh ${T}_weak 4 --steps 2 --warmup 3 --no-cpu-baseline | cut -c1-300
