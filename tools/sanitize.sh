#!/bin/bash
# compute-sanitizer passes over the smallest geometries (race detection / sanitizers).  Run on a GPU
# box:  bash tools/sanitize.sh [tag]   (memcheck ~1 min, racecheck / synccheck a few minutes; outputs under gpurun_out/)
TAG=${1:-san}
OUT=gpurun_out
mkdir -p $OUT
SEL='test_core_edge_shapes or test_add_layer_norm_against_torch or test_residual_bias_add_against_torch or test_deactivated_is_plain_attention'
for tool in memcheck synccheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 3 python -m pytest tests -m gpu -x -q -k "$SEL" \
      > $OUT/${TAG}_${tool}.log 2>&1
  echo "$tool rc=$?" | tee -a $OUT/${TAG}_${tool}.log
done
