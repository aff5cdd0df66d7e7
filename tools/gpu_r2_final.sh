#!/bin/bash
# End-of-round-2 pass on one B200: GPU test suite, smoke, default bench (+ cpu_baseline), SD1.5 lines, IP-Adapter line,
# ncu captures of the attention core and of the GEMM shapes, ncu launch list of the bench command, forward kernel breakdown,
# compute-sanitizer on the smallest geometries.   bash tools/gpu_r2_final.sh [ref]
OUT=gpurun_out; T=r2f
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > $OUT/${T}_gpu.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $OUT/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${T}_pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${T}_smoke.log 2>&1; echo "smoke rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_ -c 8 -o $OUT/${T}_attn_core -f \
    python tools/ncu_core.py > $OUT/${T}_ncu_core.log 2>&1; echo "ncu core rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:linear_tc -c 8 -o $OUT/${T}_gemm -f \
    python tools/ncu_gemm.py > $OUT/${T}_ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 14000 --csv --log-file $OUT/${T}_launches.csv \
    python bench.py --steps 1 --warmup 0 --denoise-steps 2 --no-graphs --no-e2e --no-cpu-baseline > $OUT/${T}_ncu_bench.log 2>&1; echo "ncu launches rc=$?"
timeout 600 python bench.py > $OUT/${T}_bench.json 2> $OUT/${T}_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --model sd15 --atype fused_outer --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${T}_bench_sd15_outer.json 2>> $OUT/${T}_bench.err
timeout 300 python bench.py --model sd15 --atype fused_inner --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${T}_bench_sd15_inner.json 2>> $OUT/${T}_bench.err
timeout 300 python bench.py --ip-tokens 16 --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${T}_bench_ip16.json 2>> $OUT/${T}_bench.err
timeout 200 python tools/profile_unet2.py > $OUT/${T}_unet_profile.txt 2>&1
if [ "$1" = "san" ] || [ "$2" = "san" ]; then
SEL='test_core_edge_shapes or test_deactivated_is_plain_attention or test_add_layer_norm_against_torch'
for tool in memcheck racecheck synccheck; do
  ( time timeout 150 compute-sanitizer --tool $tool --error-exitcode 3 python -m pytest tests -m gpu -x -q -k "$SEL" ) > $OUT/${T}_san_${tool}.log 2>&1
  echo "$tool rc=$?" | tee -a $OUT/${T}_san_${tool}.log
done
fi
if [ "$1" = "ref" ]; then
  timeout 600 python bench.py --impl reference --steps 4 --warmup 1 > $OUT/${T}_bench_reference.json 2>> $OUT/${T}_bench.err; echo "reference rc=$?"
fi
tail -3 $OUT/${T}_pytest_gpu.log; tail -2 $OUT/${T}_smoke.log; cut -c1-300 $OUT/${T}_bench.json; cut -c1-200 $OUT/${T}_bench_sd15_outer.json; cut -c1-200 $OUT/${T}_bench_sd15_inner.json; cut -c1-200 $OUT/${T}_bench_ip16.json; tail -3 $OUT/${T}_bench.err
