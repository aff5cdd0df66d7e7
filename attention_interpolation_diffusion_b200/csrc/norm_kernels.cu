// HBM-bound glue around the interpolated-attention path ("rest of the
// BasicTransformerBlock around the path"): the residual add + LayerNorm that feeds every attention / feed-forward
// call, and the channels-last GroupNorm (+ SiLU) in front of the transformer stack and inside the ResNet blocks.
// The reference leaves these to diffusers -> PyTorch [ext]: LayerNorm and the residual add as two kernels, GroupNorm
// as NHWC->NCHW copy + moments + normalise + SiLU + NCHW->NHWC copy (profiles/r1_unet_forward_kernels_b45e619.txt:
// 27 of 83 ms per SDXL forward).  Here each is one or two streaming passes with 16-byte accesses along the channel
// dimension; all reductions have a fixed order (no floating-point atomics), so results are bit-reproducible.
//
//   add_layer_norm_kernel : x_new = x + delta (optional), h = LN(x_new) * gamma + beta.   One warp per row, the row
//                           stays in registers (exact two-pass mean / variance).   bytes: 2*(2|1 reads + 2|1 writes) / element
//   gn_stats_kernel       : per (frame, pixel chunk, group) partial (mean, M2) of an NHWC tensor (+ optional per-(n,c)
//                           bias added on load: the ResNet time-embedding add).         bytes: 2 / element (one read)
//   gn_apply_kernel       : merges the partials (Chan), y = (x - mean) * rstd * gamma + beta, optional SiLU.
//                                                                                      bytes: 4 / element (read + write)
#include "paid_common.cuh"

namespace paid {
namespace {

constexpr int kLnWarps = 8;
constexpr int kLnMaxVec = 8;   // 16-byte vectors per lane: rows of up to 32 * 8 * 8 = 2048 channels

template <typename T>
__device__ __forceinline__ void unpack8(const uint4& v, float* f) {
  const T* p = reinterpret_cast<const T*>(&v);
#pragma unroll
  for (int e = 0; e < 8; ++e) f[e] = to_f32(p[e]);
}
template <typename T>
__device__ __forceinline__ uint4 pack8(const float* f) {
  return make_uint4(pack2<T>(f[0], f[1]), pack2<T>(f[2], f[3]), pack2<T>(f[4], f[5]), pack2<T>(f[6], f[7]));
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <typename T>
__global__ void __launch_bounds__(kLnWarps * 32)
add_layer_norm_kernel(const T* x, const T* delta, const T* __restrict__ gamma, const T* __restrict__ beta, T* x_out,
                      T* __restrict__ h_out, long long rows, int C, float eps) {   // x_out may alias x or delta
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * kLnWarps + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int nvec = C >> 3;
  const T* xr = x + row * C;
  float v[kLnMaxVec][8];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < kLnMaxVec; ++i) {
    const int j = lane + 32 * i;
    if (j < nvec) {
      unpack8<T>(*reinterpret_cast<const uint4*>(xr + j * 8), v[i]);
      if (delta) {
        float d[8];
        unpack8<T>(*reinterpret_cast<const uint4*>(delta + row * C + j * 8), d);
#pragma unroll
        for (int e = 0; e < 8; ++e) v[i][e] = to_f32(from_f32<T>(v[i][e] + d[e]));  // the residual stream is stored in T
        *reinterpret_cast<uint4*>(x_out + row * C + j * 8) = pack8<T>(v[i]);
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) sum += v[i][e];
    }
  }
  const float mean = warp_sum(sum) / (float)C;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < kLnMaxVec; ++i)
    if (lane + 32 * i < nvec) {
#pragma unroll
      for (int e = 0; e < 8; ++e) { const float c = v[i][e] - mean; sq += c * c; }
    }
  const float rstd = rsqrtf(warp_sum(sq) / (float)C + eps);
#pragma unroll
  for (int i = 0; i < kLnMaxVec; ++i) {
    const int j = lane + 32 * i;
    if (j < nvec) {
      float g[8], b[8], o[8];
      unpack8<T>(*reinterpret_cast<const uint4*>(gamma + j * 8), g);
      unpack8<T>(*reinterpret_cast<const uint4*>(beta + j * 8), b);
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = (v[i][e] - mean) * rstd * g[e] + b[e];
      *reinterpret_cast<uint4*>(h_out + row * C + j * 8) = pack8<T>(o);
    }
  }
}

// ---- GroupNorm over an NHWC tensor --------------------------------------------------------------------------------
// Thread layout of both kernels: V = C / 8 vector columns, R pixel rows in flight; thread (r, col) owns channels
// [8 col, 8 col + 8) of pixels p0 + r, p0 + r + R, ...   A group (C / groups channels, 10 ... 80 here) is not aligned to
// the 8-channel vectors, so statistics are kept per channel and folded into groups in shared memory.
struct GnGeometry {
  int V, R, threads, chunks_stats, chunks_apply;   // threads = V * R rounded up to whole warps (the rest idle in the loops)
};

__host__ inline GnGeometry gn_geometry(int N, long long HW, int C) {
  GnGeometry g;
  g.V = C / 8;
  g.R = g.V >= 384 ? 1 : 384 / g.V;
  if ((long long)g.R > HW) g.R = (int)HW;
  g.threads = (g.V * g.R + 31) & ~31;
  const long long max_chunks = (HW + g.R - 1) / g.R;
  const long long frame_bytes = HW * C * 2;
  // at least two CTAs per SM over the whole grid, and no CTA streaming much less than bytes_per_cta
  auto pick = [&](long long bytes_per_cta, long long cap) {
    long long c = (2 * 148 + N - 1) / N;
    if (frame_bytes / bytes_per_cta > c) c = frame_bytes / bytes_per_cta;
    if (c > cap) c = cap;
    if (c > max_chunks) c = max_chunks;
    return (int)(c < 1 ? 1 : c);
  };
  g.chunks_stats = pick(128 << 10, 64);    // <= 64 partials per (frame, group): merged by two lane-strided passes
  g.chunks_apply = pick(64 << 10, 65535);
  return g;
}

template <typename T>
__global__ void gn_stats_kernel(const T* __restrict__ x, const T* __restrict__ pre_bias, float2* __restrict__ partial,
                                long long HW, int C, int groups, int R, int chunks) {
  extern __shared__ float sh[];  // [2][R][C] per-channel sums and sums of squares
  const int V = C >> 3, tid = threadIdx.x, col = tid % V, r = tid / V;
  const bool active = r < R;   // the block is padded to whole warps
  const int n = blockIdx.y, chunk = blockIdx.x;
  const long long P = (HW + chunks - 1) / chunks;
  const long long p0 = chunk * P, p1 = active ? (p0 + P < HW ? p0 + P : HW) : 0;
  float pb[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (pre_bias && active) unpack8<T>(*reinterpret_cast<const uint4*>(pre_bias + (long long)n * C + col * 8), pb);
  float s[8], ss[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) s[e] = ss[e] = 0.f;
  const T* base = x + ((long long)n * HW) * C + col * 8;
  long long p = p0 + r;
  for (; p + 3LL * R < p1; p += 4LL * R) {  // four independent 16-byte loads in flight per thread
    uint4 raw[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) raw[u] = *reinterpret_cast<const uint4*>(base + (p + (long long)u * R) * C);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float f[8];
      unpack8<T>(raw[u], f);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float val = pre_bias ? to_f32(from_f32<T>(f[e] + pb[e])) : f[e];
        s[e] += val; ss[e] = fmaf(val, val, ss[e]);
      }
    }
  }
  for (; p < p1; p += R) {
    float f[8];
    unpack8<T>(*reinterpret_cast<const uint4*>(base + p * C), f);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float val = pre_bias ? to_f32(from_f32<T>(f[e] + pb[e])) : f[e];
      s[e] += val; ss[e] = fmaf(val, val, ss[e]);
    }
  }
  float* sh_s = sh;
  float* sh_ss = sh + R * C;
  if (active) {
#pragma unroll
    for (int e = 0; e < 8; ++e) { sh_s[r * C + col * 8 + e] = s[e]; sh_ss[r * C + col * 8 + e] = ss[e]; }
  }
  __syncthreads();
  // fold channels into groups in a fixed order: one thread sums its group's R x gs entries
  const long long q1 = p0 + P < HW ? p0 + P : HW;
  for (int g = tid; g < groups; g += blockDim.x) {
    const int gs = C / groups;
    float S = 0.f, SS = 0.f;
    for (int rr = 0; rr < R; ++rr)
      for (int c = g * gs; c < (g + 1) * gs; ++c) { S += sh_s[rr * C + c]; SS += sh_ss[rr * C + c]; }
    const float cnt = (float)((q1 > p0 ? q1 - p0 : 0) * gs);
    const float mean = cnt > 0.f ? S / cnt : 0.f;
    const float m2 = fmaxf(SS - S * mean, 0.f);
    partial[((long long)n * chunks + chunk) * groups + g] = make_float2(mean, m2);
  }
}

template <typename T>
__global__ void gn_apply_kernel(const T* __restrict__ x, const T* __restrict__ pre_bias, const float2* __restrict__ partial,
                                const T* __restrict__ gamma, const T* __restrict__ beta, T* __restrict__ y, long long HW,
                                int C, int groups, int R, int chunks_stats, int chunks, float eps, int silu) {
  __shared__ float sh_mean[64], sh_rstd[64];
  const int V = C >> 3, tid = threadIdx.x, col = tid % V, r = tid / V;
  const bool active = r < R;   // the block is padded to whole warps (every warp takes part in the merge below)
  const int n = blockIdx.y, chunk = blockIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  // merge the per-chunk (mean, M2) partials of this frame: warp per group, lanes over chunks (chunks_stats <= 64)
  {
    const int gs = C / groups;
    const long long Ps = (HW + chunks_stats - 1) / chunks_stats;
    for (int g = warp; g < groups; g += nwarps) {
      float cw[2], mw[2], m2w[2];
      float wsum = 0.f;
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int c = lane + 32 * u;
        cw[u] = 0.f; mw[u] = 0.f; m2w[u] = 0.f;
        if (c < chunks_stats) {
          const long long q0 = c * Ps, q1 = q0 + Ps < HW ? q0 + Ps : HW;
          const float2 pm = partial[((long long)n * chunks_stats + c) * groups + g];
          cw[u] = (float)((q1 > q0 ? q1 - q0 : 0) * gs); mw[u] = pm.x; m2w[u] = pm.y;
          wsum += cw[u] * pm.x;
        }
      }
      const float total = (float)(HW * gs);
      const float mean = warp_sum(wsum) / total;
      float m2 = 0.f;
#pragma unroll
      for (int u = 0; u < 2; ++u) { const float dlt = mw[u] - mean; m2 += m2w[u] + cw[u] * dlt * dlt; }
      m2 = warp_sum(m2);
      if (lane == 0) { sh_mean[g] = mean; sh_rstd[g] = rsqrtf(m2 / total + eps); }
    }
  }
  __syncthreads();
  if (!active) return;
  // per-channel scale / shift of this thread's 8 channels:  y = x * a + b
  float a[8], b[8], pb[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  {
    const int gs = C / groups;
    float gm[8], bt[8];
    unpack8<T>(*reinterpret_cast<const uint4*>(gamma + col * 8), gm);
    unpack8<T>(*reinterpret_cast<const uint4*>(beta + col * 8), bt);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int g = (col * 8 + e) / gs;
      a[e] = sh_rstd[g] * gm[e];
      b[e] = bt[e] - sh_mean[g] * a[e];
    }
    if (pre_bias) unpack8<T>(*reinterpret_cast<const uint4*>(pre_bias + (long long)n * C + col * 8), pb);
  }
  const long long P = (HW + chunks - 1) / chunks;
  const long long p0 = chunk * P, p1 = p0 + P < HW ? p0 + P : HW;
  const T* base = x + ((long long)n * HW) * C + col * 8;
  T* obase = y + ((long long)n * HW) * C + col * 8;
  auto transform = [&](const uint4& raw) {
    float f[8], o[8];
    unpack8<T>(raw, f);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float val = pre_bias ? to_f32(from_f32<T>(f[e] + pb[e])) : f[e];
      float t = fmaf(val, a[e], b[e]);
      if (silu) t = __fdividef(t, 1.f + __expf(-t));
      o[e] = t;
    }
    return pack8<T>(o);
  };
  long long p = p0 + r;
  for (; p + 3LL * R < p1; p += 4LL * R) {
    uint4 raw[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) raw[u] = *reinterpret_cast<const uint4*>(base + (p + (long long)u * R) * C);
#pragma unroll
    for (int u = 0; u < 4; ++u) *reinterpret_cast<uint4*>(obase + (p + (long long)u * R) * C) = transform(raw[u]);
  }
  for (; p < p1; p += R) *reinterpret_cast<uint4*>(obase + p * C) = transform(*reinterpret_cast<const uint4*>(base + p * C));
}

template <typename T>
int launch_gn_t(const void* x, const void* pre_bias, const void* gamma, const void* beta, void* y, float* ws, int N,
                long long HW, int C, int groups, float eps, int silu, cudaStream_t stream) {
  const GnGeometry g = gn_geometry(N, HW, C);
  const size_t smem = (size_t)2 * g.R * C * sizeof(float);
  gn_stats_kernel<T><<<dim3(g.chunks_stats, N), g.threads, smem, stream>>>((const T*)x, (const T*)pre_bias, (float2*)ws, HW, C,
                                                                          groups, g.R, g.chunks_stats);
  PAID_LAUNCH_CHECK("gn_stats_kernel");
  gn_apply_kernel<T><<<dim3(g.chunks_apply, N), g.threads, 0, stream>>>((const T*)x, (const T*)pre_bias, (const float2*)ws,
                                                                       (const T*)gamma, (const T*)beta, (T*)y, HW, C, groups,
                                                                       g.R, g.chunks_stats, g.chunks_apply, eps, silu);
  PAID_LAUNCH_CHECK("gn_apply_kernel");
  return PAID_OK;
}

}  // namespace

unsigned long long group_norm_workspace_bytes(int N, long long HW, int C, int groups) {
  if (N <= 0 || HW <= 0 || C <= 0 || groups <= 0) return 0;
  const GnGeometry g = gn_geometry(N, HW, C);
  return (unsigned long long)N * g.chunks_stats * groups * sizeof(float2);
}

bool group_norm_supported(int C, int groups) {
  // 16-byte channel vectors; the per-channel staging of the statistics kernel must fit the default 48 KB of shared memory
  return C % 8 == 0 && groups > 0 && groups <= 64 && C % groups == 0 && C / 8 <= 512;
}

int launch_group_norm_nhwc(const void* x, const void* pre_bias, const void* gamma, const void* beta, void* y, float* ws, int N,
                           long long HW, int C, int groups, float eps, int silu, int dtype, cudaStream_t stream) {
  return dtype == PAID_F16
             ? launch_gn_t<__half>(x, pre_bias, gamma, beta, y, ws, N, HW, C, groups, eps, silu, stream)
             : launch_gn_t<__nv_bfloat16>(x, pre_bias, gamma, beta, y, ws, N, HW, C, groups, eps, silu, stream);
}

bool add_layer_norm_supported(int C) { return C % 8 == 0 && C / 8 <= 32 * kLnMaxVec; }

int launch_add_layer_norm(const void* x, const void* delta, const void* gamma, const void* beta, void* x_out, void* h_out,
                          long long rows, int C, float eps, int dtype, cudaStream_t stream) {
  const unsigned blocks = (unsigned)((rows + kLnWarps - 1) / kLnWarps);
  if (dtype == PAID_F16)
    add_layer_norm_kernel<__half><<<blocks, kLnWarps * 32, 0, stream>>>((const __half*)x, (const __half*)delta, (const __half*)gamma,
                                                                        (const __half*)beta, (__half*)x_out, (__half*)h_out, rows, C, eps);
  else
    add_layer_norm_kernel<__nv_bfloat16><<<blocks, kLnWarps * 32, 0, stream>>>(
        (const __nv_bfloat16*)x, (const __nv_bfloat16*)delta, (const __nv_bfloat16*)gamma, (const __nv_bfloat16*)beta,
        (__nv_bfloat16*)x_out, (__nv_bfloat16*)h_out, rows, C, eps);
  PAID_LAUNCH_CHECK("add_layer_norm_kernel");
  return PAID_OK;
}

}  // namespace paid
