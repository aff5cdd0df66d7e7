#!/bin/bash
# build an experimental variant of the library: tools/build_variant.sh NAME "-DPAID_ABLATE=1 ..."
set -e
cd "$(dirname "$0")/.."
C=attention_interpolation_diffusion_b200/csrc
mkdir -p gpurun_scratch
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -shared $2 \
  $C/paid_api.cu $C/generic_kernels.cu $C/tma_util.cu $C/norm_kernels.cu $C/gemm_tc.cu $C/attn_tc.cu $C/attn_dw.cu -o gpurun_scratch/libpaid_$1.so
