#!/bin/bash
# round-2 GPU session y: 32-byte epilogue stores: GPU test suite, default bench
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
( timeout 900 python -m pytest tests -m gpu -x -q ) > $OUT/r2y_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 6 $OUT/r2y_pytest_gpu.log
timeout 400 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/r2y_bench.json 2> $OUT/r2y_bench.err; echo "bench rc=$?"; cut -c1-260 $OUT/r2y_bench.json; tail -3 $OUT/r2y_bench.err
