#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
( time timeout 400 python -m pytest tests -m gpu -x -q ) > $OUT/r2e_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/r2e_pytest_gpu.log
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/r2e_bench_dw.json 2> $OUT/r2e_bench.err
PAID_ATTN_DW=0 timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/r2e_bench_onewg.json 2>> $OUT/r2e_bench.err
tail -5 $OUT/r2e_pytest_gpu.log; cut -c1-400 $OUT/r2e_bench_dw.json; echo; cut -c1-400 $OUT/r2e_bench_onewg.json; tail -3 $OUT/r2e_bench.err
