// Issue-rate microbenchmark of the instructions of the softmax inner loop on sm_100a (one number per instruction mix):
// cycles per warp-instruction per SM sub-partition, 4 warps per sub-partition, 8 independent chains per thread.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_scratch/ubench_pipes tools/ubench_pipes.cu && gpurun_scratch/ubench_pipes
// Used to decide what limits the attention core at head_dim 64: which pipe MUFU.EX2, F2FP.PACK_AB,
// FFMA2 / FADD2, FFMA (register and immediate forms), IMAD, FMNMX, LOP3 / SHF / IADD3 occupy and for how long.
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

constexpr int ITERS = 2048;

#define KERNEL(name, BODY)                                                              \
  __global__ void name(float* out, long long* cyc, float seed) {                         \
    float r0 = seed + threadIdx.x, r1 = r0 + 1, r2 = r0 + 2, r3 = r0 + 3, r4 = r0 + 4, r5 = r0 + 5, r6 = r0 + 6, r7 = r0 + 7; \
    unsigned u0 = threadIdx.x, u1 = u0 + 1, u2 = u0 + 2, u3 = u0 + 3;                     \
    unsigned long long d0 = threadIdx.x, d1 = d0 + 7, d2 = d0 + 9, d3 = d0 + 11;          \
    __syncthreads();                                                                      \
    const long long t0 = clock64();                                                       \
    _Pragma("unroll 1") for (int i = 0; i < ITERS; ++i) { BODY }                          \
    const long long t1 = clock64();                                                       \
    out[blockIdx.x * blockDim.x + threadIdx.x] = r0 + r1 + r2 + r3 + r4 + r5 + r6 + r7 + u0 + u1 + u2 + u3 + (float)(d0 + d1 + d2 + d3); \
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;                                      \
  }

#define EX2(r) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(r));
#define CVT(u, a, b) asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(u) : "f"(a), "f"(b));
#define FFMA(r) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(r) : "f"(seed), "f"(r7));
#define FFMAI(r) asm volatile("fma.rn.f32 %0, %0, %1, 0f3F000000;" : "+f"(r) : "f"(seed));
#define FFMA2(d) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(d) : "l"(d3));
#define FADD2(d) asm volatile("add.f32x2 %0, %0, %1;" : "+l"(d) : "l"(d3));
#define FMNMX(r) asm volatile("max.f32 %0, %0, %1;" : "+f"(r) : "f"(seed));
#define IMAD(u) asm volatile("mad.lo.u32 %0, %0, 8388608, %1;" : "+r"(u) : "r"(u3));
#define IADD(u) asm volatile("add.u32 %0, %0, 4096;" : "+r"(u));
#define SHF(u) asm volatile("shr.u32 %0, %0, 13;" : "+r"(u));
#define LOP(u) asm volatile("lop3.b32 %0, %0, %1, 0xFFFF0000, 0xE4;" : "+r"(u) : "r"(u3));

KERNEL(k_ex2, EX2(r0) EX2(r1) EX2(r2) EX2(r3) EX2(r4) EX2(r5) EX2(r6) EX2(r7))                        // 8 MUFU
KERNEL(k_cvt, CVT(u0, r0, r1) CVT(u1, r2, r3) CVT(u2, r4, r5) CVT(u3, r6, r7) CVT(u0, r1, r2) CVT(u1, r3, r4) CVT(u2, r5, r6) CVT(u3, r7, r0))  // 8 F2FP
KERNEL(k_ex2_cvt, EX2(r0) EX2(r1) EX2(r2) EX2(r3) CVT(u0, r4, r5) CVT(u1, r6, r7) CVT(u2, r5, r6) CVT(u3, r7, r4))  // 4 + 4
KERNEL(k_ffma, FFMA(r0) FFMA(r1) FFMA(r2) FFMA(r3) FFMA(r4) FFMA(r5) FFMA(r6) FFMA(r0))
KERNEL(k_ffma_imm, FFMAI(r0) FFMAI(r1) FFMAI(r2) FFMAI(r3) FFMAI(r4) FFMAI(r5) FFMAI(r6) FFMAI(r7))
KERNEL(k_ffma2, FFMA2(d0) FFMA2(d1) FFMA2(d2) FFMA2(d0) FFMA2(d1) FFMA2(d2) FFMA2(d0) FFMA2(d1))
KERNEL(k_fadd2, FADD2(d0) FADD2(d1) FADD2(d2) FADD2(d0) FADD2(d1) FADD2(d2) FADD2(d0) FADD2(d1))
KERNEL(k_fmnmx, FMNMX(r0) FMNMX(r1) FMNMX(r2) FMNMX(r3) FMNMX(r4) FMNMX(r5) FMNMX(r6) FMNMX(r7))
KERNEL(k_imad, IMAD(u0) IMAD(u1) IMAD(u2) IMAD(u0) IMAD(u1) IMAD(u2) IMAD(u0) IMAD(u1))
KERNEL(k_alu, IADD(u0) SHF(u1) LOP(u2) IADD(u1) SHF(u2) LOP(u0) IADD(u2) SHF(u0))                     // 8 ALU-pipe integer ops
KERNEL(k_ex2_ffma2, EX2(r0) EX2(r1) EX2(r2) EX2(r3) FFMA2(d0) FFMA2(d1) FFMA2(d2) FFMA2(d0))           // 4 MUFU + 4 FFMA2
KERNEL(k_ex2_alu, EX2(r0) EX2(r1) EX2(r2) EX2(r3) IADD(u0) SHF(u1) LOP(u2) IADD(u1))                   // 4 MUFU + 4 ALU
KERNEL(k_ffma2_alu, FFMA2(d0) FFMA2(d1) FFMA2(d2) FFMA2(d0) IADD(u0) SHF(u1) LOP(u2) IADD(u1))         // 4 FFMA2 + 4 ALU
KERNEL(k_ffma2_fmnmx, FFMA2(d0) FFMA2(d1) FFMA2(d2) FFMA2(d0) FMNMX(r0) FMNMX(r1) FMNMX(r2) FMNMX(r3))
KERNEL(k_cvt_fmnmx, CVT(u0, r4, r5) CVT(u1, r6, r7) CVT(u2, r5, r6) CVT(u3, r7, r4) FMNMX(r0) FMNMX(r1) FMNMX(r2) FMNMX(r3))

template <typename K>
void run(const char* name, K kern, int threads, float* out, long long* cyc) {
  kern<<<148, threads>>>(out, cyc, 0.5f);
  kern<<<148, threads>>>(out, cyc, 0.5f);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < 148; ++i) avg += h[i];
  avg /= 148;
  const int warps_per_smsp = threads / 128;
  printf("{\"mix\": \"%s\", \"warps_per_smsp\": %d, \"cycles_per_warp_instr_per_smsp\": %.2f}\n", name, warps_per_smsp,
         avg / ITERS / 8 / warps_per_smsp);
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * sizeof(float));
  cudaMalloc(&cyc, 148 * sizeof(long long));
  for (int threads : {128, 512}) {
    run("8 MUFU.EX2", k_ex2, threads, out, cyc);
    run("8 F2FP.F16.F32.PACK_AB", k_cvt, threads, out, cyc);
    run("4 MUFU.EX2 + 4 F2FP", k_ex2_cvt, threads, out, cyc);
    run("8 FFMA (3 registers)", k_ffma, threads, out, cyc);
    run("8 FFMA (immediate addend)", k_ffma_imm, threads, out, cyc);
    run("8 FFMA2", k_ffma2, threads, out, cyc);
    run("8 FADD2", k_fadd2, threads, out, cyc);
    run("8 FMNMX", k_fmnmx, threads, out, cyc);
    run("8 IMAD", k_imad, threads, out, cyc);
    run("8 IADD/SHF/LOP3", k_alu, threads, out, cyc);
    run("4 MUFU.EX2 + 4 FFMA2", k_ex2_ffma2, threads, out, cyc);
    run("4 MUFU.EX2 + 4 IADD/SHF/LOP3", k_ex2_alu, threads, out, cyc);
    run("4 FFMA2 + 4 IADD/SHF/LOP3", k_ffma2_alu, threads, out, cyc);
    run("4 FFMA2 + 4 FMNMX", k_ffma2_fmnmx, threads, out, cyc);
    run("4 F2FP + 4 FMNMX", k_cvt_fmnmx, threads, out, cyc);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
