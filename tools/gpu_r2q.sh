#!/bin/bash
# round-2 GPU session q: dual-Q-block attn_dw + FMA-pipe exponentials.  Variants built by tools/build_variant.sh.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 300 python tools/try_dw.py ) > gpurun_out/r2q_try_dw_default.log 2>&1; echo "default rc=$?"
for v in p0 p6 p8; do
  ( PAID_LIB_PATH=$PWD/gpurun_scratch/libpaid_$v.so timeout 300 python tools/try_dw.py ) > gpurun_out/r2q_try_dw_$v.log 2>&1; echo "$v rc=$?"
done
( PAID_LIB_PATH=$PWD/gpurun_scratch/libpaid_trace.so timeout 120 python tools/trace_dw.py ) > gpurun_out/r2q_trace.log 2>&1; echo "trace rc=$?"
tail -n 13 gpurun_out/r2q_try_dw_default.log
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2q_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 5 gpurun_out/r2q_pytest_gpu.log
