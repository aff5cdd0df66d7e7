#!/bin/bash
# One GPU-box pass: parity tests, smoke, ncu captures, bench (SDXL own arm, SD1.5), forward kernel breakdown, ncu launch
# list.  Usage (repo root, under gpurun):  bash tools/gpu_validate.sh <tag> [ref]
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > $OUT/${TAG}_gpu.txt 2>&1
# the chunked wide-head kernel first, in its own process: if it fails, the rest of the pass runs with head_dim > 64 on
# the generic kernels so that the remaining measurements stay meaningful
( timeout 300 python -m pytest tests -m gpu -x -q -k "sd15 or wide or golden or edge" ) > $OUT/${TAG}_pytest_wide.log 2>&1
WIDE_RC=$?
echo "wide rc=$WIDE_RC" >> $OUT/${TAG}_pytest_wide.log
if [ $WIDE_RC -ne 0 ]; then export PAID_ATTN_MAX_TC_HEAD_DIM=64; fi
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $OUT/${TAG}_pytest_gpu.log 2>&1
echo "pytest rc=$? (PAID_ATTN_MAX_TC_HEAD_DIM=$PAID_ATTN_MAX_TC_HEAD_DIM)" >> $OUT/${TAG}_pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1
timeout 600 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
timeout 300 python bench.py --model sd15 --atype fused_outer --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_sd15_outer.json 2>> $OUT/${TAG}_bench.err
timeout 300 python bench.py --model sd15 --atype fused_inner --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/${TAG}_bench_sd15_inner.json 2>> $OUT/${TAG}_bench.err
if [ "$2" = "ref" ]; then
  timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > $OUT/${TAG}_bench_reference.json 2>> $OUT/${TAG}_bench.err
fi
timeout 200 python tools/profile_unet2.py > $OUT/${TAG}_unet_profile.txt 2>&1
timeout 300 ncu --set full --clock-control none -k regex:"gn_|layer_norm|geglu|residual_bias" -c 16 -o $OUT/${TAG}_glue -f \
    python tools/ncu_glue.py > $OUT/${TAG}_ncu_glue.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_tc -c 8 -o $OUT/${TAG}_attn_core -f \
    python tools/ncu_core.py > $OUT/${TAG}_ncu_core.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 5600 -c 5000 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 1 --denoise-steps 2 --no-graphs --no-e2e --no-cpu-baseline > $OUT/${TAG}_ncu_bench.log 2>&1
tail -3 $OUT/${TAG}_pytest_wide.log; tail -3 $OUT/${TAG}_pytest_gpu.log; tail -2 $OUT/${TAG}_smoke.log; cut -c1-400 $OUT/${TAG}_bench.json; cut -c1-200 $OUT/${TAG}_bench_sd15_outer.json
