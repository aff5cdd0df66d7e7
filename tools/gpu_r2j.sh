#!/bin/bash
# Two-GPU pass: the merged-pass test on GPU 0, the multi-rank NCCL parity test at world 2, weak-scaling and configs[3] bench lines at 2 GPUs.
OUT=gpurun_out; T=${1:-r2j}
mkdir -p $OUT
( timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -q -k "merged_passes or graph_replay" ) > $OUT/${T}_pytest_merged.log 2>&1; echo "pytest merged rc=$?"; tail -2 $OUT/${T}_pytest_merged.log
( time timeout 900 python -m pytest tests/test_multirank_gpu.py -m gpu -q ) > $OUT/${T}_pytest_multirank.log 2>&1; echo "pytest multirank rc=$?" | tee -a $OUT/${T}_pytest_multirank.log; tail -4 $OUT/${T}_pytest_multirank.log
bash tools/gpu_scale.sh ${T}_weak 2 --steps 2 --warmup 3 --no-cpu-baseline | cut -c1-700
bash tools/gpu_scale.sh ${T}_c3 2 --frames 16 --steps 1 --warmup 3 --no-cpu-baseline | cut -c1-700
