#!/bin/bash
# round-2 GPU session t: softmax-loop variants of attn_dw (timing only), default bench + UNet kernel profile at HEAD
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
( TRY_DW_TIMING_ONLY=1 timeout 200 python tools/try_dw.py ) > $OUT/r2t_try_dw_default.log 2>&1; echo "default rc=$?"
for v in nv sumchk nvsum; do
  ( TRY_DW_TIMING_ONLY=1 PAID_LIB_PATH=$PWD/gpurun_scratch/libpaid_$v.so timeout 200 python tools/try_dw.py ) > $OUT/r2t_try_dw_$v.log 2>&1; echo "$v rc=$?"
done
grep -h '"plain"' $OUT/r2t_try_dw_*.log | cut -c1-170
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/r2t_bench.json 2> $OUT/r2t_bench.err; echo "bench rc=$?"
cut -c1-300 $OUT/r2t_bench.json; tail -3 $OUT/r2t_bench.err
timeout 200 python tools/profile_unet.py > $OUT/r2t_unet_profile.txt 2>&1; echo "profile rc=$?"; head -12 $OUT/r2t_unet_profile.txt
