// Shared host/device helpers of libpaid_attn (internal; the public surface is include/paid_attn.h).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "../../include/paid_attn.h"

namespace paid {

// ---- error reporting (thread-local message, see paid_attn_last_error) ------------------------
char* error_buffer();              // 512 bytes, thread-local
const char** last_kernel_slot();   // thread-local
int fail(int status, const char* fmt, ...);
std::atomic<uint64_t>& launch_counter();

#define PAID_CUDA_CHECK(expr)                                                                  \
  do {                                                                                         \
    cudaError_t e__ = (expr);                                                                  \
    if (e__ != cudaSuccess)                                                                    \
      return ::paid::fail(PAID_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), \
                          __FILE__, __LINE__);                                                 \
  } while (0)

#define PAID_LAUNCH_CHECK(name)                                                          \
  do {                                                                                   \
    ::paid::launch_counter().fetch_add(1, std::memory_order_relaxed);                    \
    cudaError_t e__ = cudaGetLastError();                                                \
    if (e__ != cudaSuccess)                                                              \
      return ::paid::fail(PAID_ECUDA, "launch of %s failed: %s", name, cudaGetErrorString(e__)); \
  } while (0)

// ---- dtype helpers --------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// pack two floats into one 32-bit word of T (lo = a, hi = b)
template <typename T> __device__ __forceinline__ uint32_t pack2(float a, float b);
template <> __device__ __forceinline__ uint32_t pack2<__half>(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
template <> __device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

constexpr float kLog2e = 1.4426950408889634f;

#ifdef __CUDACC__
// exact-form GELU 0.5 g (1 + erf(g / sqrt 2)) with erf from Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7, far below the
// 16-bit output rounding): one MUFU.RCP, one MUFU.EX2 and 7 FMA-pipe instructions instead of erff's ~30.
__device__ __forceinline__ float gelu_erf(float g) {
  const float z = fabsf(g) * 0.70710678118654752f;
  const float t = __fdividef(1.f, fmaf(0.3275911f, z, 1.f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float erfc_z = p * t * exp2f(-z * z * 1.4426950408889634f);   // 1 - erf(|g| / sqrt 2)
  const float half_g = 0.5f * g;
  // g >= 0: 0.5 g (2 - erfc) ;  g < 0: 0.5 g erfc
  return g >= 0.f ? fmaf(-half_g, erfc_z, g) : half_g * erfc_z;
}
#endif

// ---- per-frame slot plan of the interpolated attention -----------------------------------------
// The attention of frame n is a combination of up to three partial (flash-style) attentions:
//   slot 0: the frame's own K/V           (PLAIN, or any fused mode)
//   slot 1: "A": begin endpoint (OUTER) or the per-frame lerped endpoint K/V (INNER)
//   slot 2: "B": end endpoint (OUTER)
//   out = wA * merge(slot0, slot1) + wB * merge(slot0, slot2)      
// A slot is skipped when its weight is exactly 0, or when it would duplicate slot 0 (an endpoint
// frame attending to its own K/V twice: merge(p, p) == norm(p) exactly).
struct FramePlan {
  float wA, wB;
  bool use0, use1, use2;
};

__host__ __device__ inline FramePlan make_frame_plan(int mode, int fused, int n, int begin_frame, int end_frame,
                                                     float c) {
  FramePlan p;
  if (mode == PAID_PLAIN) {
    p.wA = 1.f; p.wB = 0.f; p.use0 = true; p.use1 = false; p.use2 = false;
    return p;
  }
  p.use0 = fused != 0;
  if (mode == PAID_OUTER) {
    p.wA = 1.f - c; p.wB = c;
    p.use1 = p.wA != 0.f && !(fused && n == begin_frame);
    p.use2 = p.wB != 0.f && !(fused && n == end_frame);
  } else {  // INNER
    p.wA = 1.f; p.wB = 0.f;
    bool dup = fused && ((n == begin_frame && c == 0.f) || (n == end_frame && c == 1.f));
    p.use1 = !dup;
    p.use2 = false;
  }
  return p;
}

// merge coefficients: out = cf0*O0 + cf1*O1 + cf2*O2 with (O_s, l_s, m_s) the un-normalised partial
// sums, their denominators and their reference maxima (log2 domain); absent slots have l = 0, m = -inf.
__device__ __forceinline__ void merge_coefficients(const FramePlan& p, float m0, float l0, float m1, float l1,
                                                   float m2, float l2, float& cf0, float& cf1, float& cf2) {
  cf0 = cf1 = cf2 = 0.f;
  if (p.wA != 0.f) {
    float M = fmaxf(p.use0 ? m0 : -INFINITY, p.use1 ? m1 : -INFINITY);
    float a0 = p.use0 ? exp2f(m0 - M) : 0.f;
    float a1 = p.use1 ? exp2f(m1 - M) : 0.f;
    float inv = p.wA / (a0 * l0 + a1 * l1);
    cf0 += a0 * inv;
    cf1 = a1 * inv;
  }
  if (p.wB != 0.f) {
    float M = fmaxf(p.use0 ? m0 : -INFINITY, p.use2 ? m2 : -INFINITY);
    float b0 = p.use0 ? exp2f(m0 - M) : 0.f;
    float b2 = p.use2 ? exp2f(m2 - M) : 0.f;
    float inv = p.wB / (b0 * l0 + b2 * l2);
    cf0 += b0 * inv;
    cf2 = b2 * inv;
  }
}

// ---- arguments shared by both attention kernel families ----------------------------------------
struct CoreArgs {
  int dtype, mode, fused;
  int N, S, L, heads, head_dim;
  float scale;
  int begin_frame, end_frame;
  const void* q;
  const void* k;
  const void* v;
  // slot 1 / slot 2 sources (already resolved): pointer to a (frames?, L, C) tensor and the frame
  // stride in elements (0: one shared (L,C) matrix for every frame)
  const void* k1; const void* v1; long long stride1;
  const void* k2; const void* v2; long long stride2;
  const float* coef;
  void* out;
  // out = (accumulate ? out : 0) + out_scale * (out_frame_scale ? out_frame_scale[n] : 1) * attention
  int accumulate;
  float out_scale;
  const float* out_frame_scale;
  long long stride0;  // frame stride of k / v in elements (0: one matrix shared by all frames)
};

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-(kernel, device) attribute: set it the first time `kernel` is
// launched on the current device (thread-safe).  Also returns the SM count of the current device.
cudaError_t ensure_kernel_configured(const void* kernel, int smem_bytes, int* num_sms);

// measurement hook (paid_attn_profile_*): called by paid_api.cu immediately around the attention and GEMM launchers
// (kind: PAID_PROFILE_*; d0..d3: the shape key of include/paid_attn.h; flops: algorithmic flops of the launch)
void profile_mark_begin(cudaStream_t stream, int kind, long long d0, long long d1, long long d2, long long d3, double flops);
void profile_mark_end(cudaStream_t stream);


// kernel launchers (each returns a PaidStatus)
int launch_linear_generic(const void* x, const void* w, const void* bias, void* y, long long M, int Nout, int K,
                          int dtype, cudaStream_t stream);
int launch_linear_tc(const void* x, const void* w, const void* bias, void* y, long long M, int Nout, int K,
                     int dtype, cudaStream_t stream);
int launch_linear_geglu_tc(const void* x, const void* w, const void* bias, void* y, long long M, int D, int K, int dtype,
                           cudaStream_t stream);
bool linear_geglu_tc_supported(long long M, int D, int K);
int launch_linear_geglu_generic(const void* x, const void* w, const void* bias, void* y, long long M, int D, int K, int dtype,
                                cudaStream_t stream);
int launch_linear_tc_grouped(const void* x, const void* const* w, const void* const* bias, void* const* y, int groups,
                             long long M, int Nout, int K, int dtype, cudaStream_t stream);
bool linear_tc_supported(long long M, int Nout, int K);
int launch_attn_generic(const CoreArgs& a, cudaStream_t stream);
int launch_attn_tc(const CoreArgs& a, cudaStream_t stream);
bool attn_tc_supported(const CoreArgs& a);
// persistent dual-warpgroup kernel of the single-stream modes (PLAIN, INNER), head_dim <= 64 (attn_dw.cu)
int launch_attn_dw(const CoreArgs& a, cudaStream_t stream);
bool attn_dw_supported(const CoreArgs& a);
int launch_geglu(const void* h, void* out, long long M, int D, int dtype, cudaStream_t stream);
int launch_add_layer_norm(const void* x, const void* delta, const void* gamma, const void* beta, void* x_out, void* h_out,
                          long long rows, int C, float eps, int dtype, cudaStream_t stream);
bool add_layer_norm_supported(int C);
int launch_residual_bias_add(const void* a, const void* b, const void* bias, void* out, long long rows, int C, int dtype,
                             cudaStream_t stream);
int launch_group_norm_nhwc(const void* x, const void* pre_bias, const void* gamma, const void* beta, void* y, float* ws, int N,
                           long long HW, int C, int groups, float eps, int silu, int dtype, cudaStream_t stream);
unsigned long long group_norm_workspace_bytes(int N, long long HW, int C, int groups);
bool group_norm_supported(int C, int groups);
int launch_lerp_endpoints(const void* kb, const void* vb, const void* ke, const void* ve, const float* coef,
                          void* kx, void* vx, int N, long long LC, int dtype, cudaStream_t stream);

}  // namespace paid
