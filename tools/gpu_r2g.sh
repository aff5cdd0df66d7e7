#!/bin/bash
# Round-2 follow-up pass on one B200: GPU tests of the tree, default bench line, range-profiled ncu launch list of the bench
# command (timed steps only), compute-sanitizer passes on tools/sanitize_target.py.
OUT=gpurun_out; T=${1:-r2g}
mkdir -p $OUT
( time timeout 600 python -m pytest tests -m gpu -x -q ) > $OUT/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${T}_pytest_gpu.log
tail -4 $OUT/${T}_pytest_gpu.log
timeout 400 python bench.py --no-cpu-baseline > $OUT/${T}_bench.json 2> $OUT/${T}_bench.err; echo "bench rc=$?"
cut -c1-260 $OUT/${T}_bench.json; tail -3 $OUT/${T}_bench.err
for tool in memcheck synccheck racecheck; do
  ( time timeout 200 compute-sanitizer --tool $tool --error-exitcode 3 python tools/sanitize_target.py ) > $OUT/${T}_san_${tool}.log 2>&1
  echo "$tool rc=$?" | tee -a $OUT/${T}_san_${tool}.log
  tail -4 $OUT/${T}_san_${tool}.log | cut -c1-200
done
( time timeout 420 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/${T}_launches.csv \
    python bench.py --steps 1 --warmup 1 --denoise-steps 2 --no-graphs --no-e2e --no-cpu-baseline --ncu-range ) > $OUT/${T}_ncu_bench.log 2>&1; echo "ncu launches rc=$?"
wc -l $OUT/${T}_launches.csv; tail -3 $OUT/${T}_ncu_bench.log | cut -c1-300
