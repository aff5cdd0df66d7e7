#!/bin/bash
# One GPU-box pass: parity tests, bench (own + reference arm), ncu launch list, ncu full capture of the attention core.
# Usage (from the repo root, under gpurun):  bash tools/gpu_validate.sh <tag>
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > $OUT/${TAG}_gpu.txt 2>&1
( time timeout 600 python -m pytest tests -m gpu -x -q ) > $OUT/${TAG}_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $OUT/${TAG}_pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1
timeout 600 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > $OUT/${TAG}_bench_reference.json 2>> $OUT/${TAG}_bench.err
timeout 200 python tools/bench_layers.py > $OUT/${TAG}_layers.json 2>&1
timeout 200 python tools/profile_unet2.py > $OUT/${TAG}_unet_profile.txt 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 5600 -c 5000 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 1 --denoise-steps 2 --no-graphs --no-e2e --no-cpu-baseline > $OUT/${TAG}_ncu_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_tc -c 4 -o $OUT/${TAG}_attn_core -f \
    python tools/ncu_core.py > $OUT/${TAG}_ncu_core.log 2>&1
tail -3 $OUT/${TAG}_pytest_gpu.log; cat $OUT/${TAG}_smoke.log | tail -2; cut -c1-600 $OUT/${TAG}_bench.json
