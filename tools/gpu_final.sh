#!/bin/bash
# Short end-of-round pass: full GPU test suite, glue-kernel microbenchmark (both GroupNorm variants), IP-Adapter
# morphing bench, default bench.   bash tools/gpu_final.sh <tag>
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
( time timeout 200 python -m pytest tests -m gpu -x -q ) > $OUT/${TAG}_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $OUT/${TAG}_pytest_gpu.log
timeout 90 python tools/bench_glue.py > $OUT/${TAG}_glue_bench.jsonl 2>&1
timeout 150 python bench.py --ip-tokens 16 --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_ip16.json 2> $OUT/${TAG}_bench.err
timeout 150 python bench.py --no-cpu-baseline > $OUT/${TAG}_bench.json 2>> $OUT/${TAG}_bench.err
tail -3 $OUT/${TAG}_pytest_gpu.log; cut -c1-300 $OUT/${TAG}_bench_ip16.json; cut -c1-300 $OUT/${TAG}_bench.json; tail -3 $OUT/${TAG}_bench.err
