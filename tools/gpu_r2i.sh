#!/bin/bash
# Round-2 confirmation pass on one B200: full GPU suite, compute-sanitizer on tools/sanitize_target.py, ncu --set full capture
# of the attention core -> profiles/attn_traffic.json (so the bench line below carries roofline.traffic), default bench line.
OUT=gpurun_out; T=${1:-r2i}
mkdir -p $OUT
( time timeout 700 python -m pytest tests -m gpu -q ) > $OUT/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${T}_pytest_gpu.log
tail -4 $OUT/${T}_pytest_gpu.log
for tool in memcheck synccheck racecheck; do
  ( time timeout 200 compute-sanitizer --tool $tool --error-exitcode 3 python tools/sanitize_target.py ) > $OUT/${T}_san_${tool}.log 2>&1
  echo "$tool rc=$?" | tee -a $OUT/${T}_san_${tool}.log
  grep -E "SUMMARY|detected" $OUT/${T}_san_${tool}.log | sort | uniq -c | head -5
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_ -c 8 -o $OUT/${T}_attn_core -f \
    python tools/ncu_core.py > $OUT/${T}_ncu_core.log 2>&1; echo "ncu core rc=$?"
python tools/attn_traffic.py $OUT/${T}_attn_core.ncu-rep profiles/attn_traffic.json > /dev/null 2> $OUT/${T}_traffic.err; echo "traffic rc=$?"
cp profiles/attn_traffic.json $OUT/${T}_attn_traffic.json
timeout 600 python bench.py > $OUT/${T}_bench.json 2> $OUT/${T}_bench.err; echo "bench rc=$?"
cut -c1-260 $OUT/${T}_bench.json; tail -3 $OUT/${T}_bench.err
