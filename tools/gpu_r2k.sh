#!/bin/bash
# glue kernels with programmatic dependent launch: full GPU suite, bench line, synccheck with an explicit barrier count
OUT=gpurun_out; T=${1:-r2k}
mkdir -p $OUT
( time timeout 700 python -m pytest tests -m gpu -q ) > $OUT/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${T}_pytest_gpu.log
tail -4 $OUT/${T}_pytest_gpu.log
timeout 400 python bench.py --no-cpu-baseline > $OUT/${T}_bench.json 2> $OUT/${T}_bench.err; echo "bench rc=$?"
cut -c1-260 $OUT/${T}_bench.json; tail -3 $OUT/${T}_bench.err
( timeout 100 compute-sanitizer --tool synccheck --num-cuda-barriers 64 --error-exitcode 3 python tools/sanitize_target.py ) > $OUT/${T}_san_synccheck_nb64.log 2>&1; echo "synccheck nb64 rc=$?"
grep -E "SUMMARY|detected|Warning" $OUT/${T}_san_synccheck_nb64.log | sort | uniq -c | head -5
( timeout 100 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanitize_target.py ) > $OUT/${T}_san_memcheck.log 2>&1; echo "memcheck rc=$?"
