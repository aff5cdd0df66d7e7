#!/bin/bash
# round-2 first GPU pass: e2e drift value, SP=2 split-row softmax try, sanitizer timing on the smallest selection
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/r2a_smi.txt
( timeout 300 python -m pytest tests -m gpu -x -q -s -k "e2e" ) > $OUT/r2a_e2e.log 2>&1; echo "e2e rc=$?" >> $OUT/r2a_e2e.log
( timeout 150 python tools/try_split.py ) > $OUT/r2a_try_split.log 2>&1; echo "split rc=$?" >> $OUT/r2a_try_split.log
for tool in memcheck synccheck racecheck; do
  ( time timeout 170 compute-sanitizer --tool $tool --error-exitcode 3 python -m pytest tests -m gpu -x -q -k "test_core_edge_shapes" ) > $OUT/r2a_san_${tool}.log 2>&1
  echo "$tool rc=$?" >> $OUT/r2a_san_${tool}.log
done
tail -3 $OUT/r2a_e2e.log; tail -12 $OUT/r2a_try_split.log; for t in memcheck synccheck racecheck; do tail -6 $OUT/r2a_san_$t.log; done
