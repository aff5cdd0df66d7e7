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
#include <cstdlib>

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
// one 16-byte read-only load (a dereferenced uint4 temporary handed to unpack8 decays into eight 2-byte loads)
template <typename T>
__device__ __forceinline__ uint4 ldg16(const T* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// NV = 16-byte vectors per lane (rows of up to 256 NV channels).  All loads of a row are issued before the first
// store: x_out may alias x or delta, so the compiler must not be asked to move loads across stores.
template <typename T, int NV>
__global__ void __launch_bounds__(kLnWarps * 32)
add_layer_norm_kernel(const T* x, const T* delta, const T* __restrict__ gamma, const T* __restrict__ beta, T* x_out,
                      T* __restrict__ h_out, long long rows, int C, float eps) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * kLnWarps + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int nvec = C >> 3;
  const T* xr = x + row * C;
  uint4 xraw[NV], draw[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int j = lane + 32 * i;
    xraw[i] = j < nvec ? *reinterpret_cast<const uint4*>(xr + j * 8) : make_uint4(0, 0, 0, 0);
  }
  if (delta) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int j = lane + 32 * i;
      draw[i] = j < nvec ? *reinterpret_cast<const uint4*>(delta + row * C + j * 8) : make_uint4(0, 0, 0, 0);
    }
  }
  float v[NV][8];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int j = lane + 32 * i;
    unpack8<T>(xraw[i], v[i]);
    if (delta) {
      float d[8];
      unpack8<T>(draw[i], d);
#pragma unroll
      for (int e = 0; e < 8; ++e) v[i][e] = to_f32(from_f32<T>(v[i][e] + d[e]));  // the residual stream is stored in T
      if (j < nvec) *reinterpret_cast<uint4*>(x_out + row * C + j * 8) = pack8<T>(v[i]);
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) sum += v[i][e];   // vectors beyond the row are zero
  }
  const float mean = warp_sum(sum) / (float)C;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i)
    if (lane + 32 * i < nvec) {
#pragma unroll
      for (int e = 0; e < 8; ++e) { const float c = v[i][e] - mean; sq += c * c; }
    }
  const float rstd = rsqrtf(warp_sum(sq) / (float)C + eps);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int j = lane + 32 * i;
    if (j < nvec) {
      float g[8], b[8], o[8];
      unpack8<T>(ldg16(gamma + j * 8), g);
      unpack8<T>(ldg16(beta + j * 8), b);
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = (v[i][e] - mean) * rstd * g[e] + b[e];
      *reinterpret_cast<uint4*>(h_out + row * C + j * 8) = pack8<T>(o);
    }
  }
}

// out = a + b + bias[c]: the ResNet block's output (shortcut + second conv + the conv biases) in one pass
template <typename T>
__global__ void __launch_bounds__(256)
residual_bias_add_kernel(const T* a, const T* b, const T* __restrict__ bias, T* out, unsigned total, unsigned V) {
  // 32-bit index arithmetic (the launcher checks rows * C / 8 < 2^31)
  const unsigned stride = gridDim.x * blockDim.x;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += 4 * stride) {
    uint4 ra[4], rb[4], rc[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const unsigned k = i + u * stride;
      if (k < total) {
        ra[u] = *reinterpret_cast<const uint4*>(a + (size_t)k * 8);
        rb[u] = *reinterpret_cast<const uint4*>(b + (size_t)k * 8);
        rc[u] = ldg16(bias + (k % V) * 8);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const unsigned k = i + u * stride;
      if (k < total) {
        float fa[8], fb[8], fc[8], o[8];
        unpack8<T>(ra[u], fa);
        unpack8<T>(rb[u], fb);
        unpack8<T>(rc[u], fc);
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = fa[e] + (fb[e] + fc[e]);
        *reinterpret_cast<uint4*>(out + (size_t)k * 8) = pack8<T>(o);
      }
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
  g.R = g.V >= 256 ? 1 : 256 / g.V;
  if ((long long)g.R > HW) g.R = (int)HW;
  g.threads = (g.V * g.R + 31) & ~31;
  const long long max_chunks = (HW + g.R - 1) / g.R;
  const long long frame_bytes = HW * C * 2;
  // at least four CTAs per SM over the whole grid, and no CTA streaming much less than bytes_per_cta
  auto pick = [&](long long bytes_per_cta, long long cap) {
    long long c = (4 * 148 + N - 1) / N;
    if (frame_bytes / bytes_per_cta > c) c = frame_bytes / bytes_per_cta;
    if (c > cap) c = cap;
    if (c > max_chunks) c = max_chunks;
    return (int)(c < 1 ? 1 : c);
  };
  g.chunks_stats = pick(64 << 10, 128);    // <= 128 partials per (frame, group): merged by four lane-strided passes
  g.chunks_apply = pick(64 << 10, 65535);
  return g;
}

// U = independent 16-byte loads in flight per thread; HB = has pre_bias; MAXT / MINB = launch bounds (register budget)
template <typename T, int U, bool HB, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) gn_stats_kernel(const T* __restrict__ x, const T* __restrict__ pre_bias, float2* __restrict__ partial,
                                long long HW, int C, int groups, int R, int chunks) {
  extern __shared__ float sh[];  // [2][R][C] per-channel sums and sums of squares
  const int V = C >> 3, tid = threadIdx.x, col = tid % V, r = tid / V;
  const bool active = r < R;   // the block is padded to whole warps
  const int n = blockIdx.y, chunk = blockIdx.x;
  const long long P = (HW + chunks - 1) / chunks;
  const long long p0 = chunk * P, p1 = active ? (p0 + P < HW ? p0 + P : HW) : 0;
  float pb[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (HB && active) unpack8<T>(ldg16(pre_bias + (long long)n * C + col * 8), pb);
  float s[8], ss[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) s[e] = ss[e] = 0.f;
  const T* base = x + ((long long)n * HW) * C + col * 8;
  for (long long p = p0 + r; p < p1; p += (long long)U * R) {  // U independent 16-byte loads in flight
    uint4 raw[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long pp = p + (long long)u * R;
      raw[u] = pp < p1 ? *reinterpret_cast<const uint4*>(base + pp * C) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (p + (long long)u * R < p1) {
        float f[8];
        unpack8<T>(raw[u], f);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float val = HB ? to_f32(from_f32<T>(f[e] + pb[e])) : f[e];
          s[e] += val; ss[e] = fmaf(val, val, ss[e]);
        }
      }
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

template <typename T, int U, bool HB, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) gn_apply_kernel(const T* __restrict__ x, const T* __restrict__ pre_bias, const float2* __restrict__ partial,
                                const T* __restrict__ gamma, const T* __restrict__ beta, T* __restrict__ y, long long HW,
                                int C, int groups, int R, int chunks_stats, int chunks, float eps, int silu) {
  __shared__ float sh_mean[64], sh_rstd[64];
  const int V = C >> 3, tid = threadIdx.x, col = tid % V, r = tid / V;
  const bool active = r < R;   // the block is padded to whole warps (every warp takes part in the merge below)
  const int n = blockIdx.y, chunk = blockIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  // merge the per-chunk (mean, M2) partials of this frame: warp per group, lanes over chunks (chunks_stats <= 128)
  {
    const int gs = C / groups;
    const long long Ps = (HW + chunks_stats - 1) / chunks_stats;
    for (int g = warp; g < groups; g += nwarps) {
      float cw[4], mw[4], m2w[4];
      float wsum = 0.f;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
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
      for (int u = 0; u < 4; ++u) { const float dlt = mw[u] - mean; m2 += m2w[u] + cw[u] * dlt * dlt; }
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
    unpack8<T>(ldg16(gamma + col * 8), gm);
    unpack8<T>(ldg16(beta + col * 8), bt);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int g = (col * 8 + e) / gs;
      a[e] = sh_rstd[g] * gm[e];
      b[e] = bt[e] - sh_mean[g] * a[e];
    }
    if (HB) unpack8<T>(ldg16(pre_bias + (long long)n * C + col * 8), pb);
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
      const float val = HB ? to_f32(from_f32<T>(f[e] + pb[e])) : f[e];
      float t = fmaf(val, a[e], b[e]);
      if (silu) t = __fdividef(t, 1.f + __expf(-t));
      o[e] = t;
    }
    return pack8<T>(o);
  };
  for (long long p = p0 + r; p < p1; p += (long long)U * R) {
    uint4 raw[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long pp = p + (long long)u * R;
      raw[u] = pp < p1 ? *reinterpret_cast<const uint4*>(base + pp * C) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long pp = p + (long long)u * R;
      if (pp < p1) *reinterpret_cast<uint4*>(obase + pp * C) = transform(raw[u]);
    }
  }
}

template <typename T, int U, bool HB, int MAXT, int MINB>
int launch_gn_v(const GnGeometry& g, const void* x, const void* pre_bias, const void* gamma, const void* beta, void* y, float* ws,
                int N, long long HW, int C, int groups, float eps, int silu, cudaStream_t stream) {
  const size_t smem = (size_t)2 * g.R * C * sizeof(float);
  gn_stats_kernel<T, U, HB, MAXT, MINB><<<dim3(g.chunks_stats, N), g.threads, smem, stream>>>(
      (const T*)x, (const T*)pre_bias, (float2*)ws, HW, C, groups, g.R, g.chunks_stats);
  PAID_LAUNCH_CHECK("gn_stats_kernel");
  gn_apply_kernel<T, U, HB, MAXT, MINB><<<dim3(g.chunks_apply, N), g.threads, 0, stream>>>(
      (const T*)x, (const T*)pre_bias, (const float2*)ws, (const T*)gamma, (const T*)beta, (T*)y, HW, C, groups, g.R,
      g.chunks_stats, g.chunks_apply, eps, silu);
  PAID_LAUNCH_CHECK("gn_apply_kernel");
  return PAID_OK;
}

// Two register / occupancy trade-offs of the same kernels (PAID_GN_VARIANT forces one):
//   0: 8 loads in flight per thread, ~100 registers, 16 warps per SM
//   1: 4 loads in flight per thread, <= 64 registers, 32 warps per SM
// Measured on B200 (tools/bench_glue.py, L2 flushed, N = 7, profiles/r1_glue_bench.jsonl): variant 1 is 8-17 % faster on
// the 64x64 and 128x128 feature maps (many pixels per CTA: occupancy hides the latency), variant 0 is 10-16 % faster on
// the 32x32 maps (few pixels per CTA: everything a thread will ever load is in flight at once).
__host__ inline int gn_default_variant(long long HW) { return HW >= 4096 ? 1 : 0; }

template <typename T, bool HB>
int launch_gn_hb(const void* x, const void* pre_bias, const void* gamma, const void* beta, void* y, float* ws, int N,
                 long long HW, int C, int groups, float eps, int silu, cudaStream_t stream) {
  const GnGeometry g = gn_geometry(N, HW, C);
  const char* env = getenv("PAID_GN_VARIANT");
  const int variant = env ? atoi(env) : gn_default_variant(HW);
  const bool small = g.threads <= 256;
#define PAID_GN_GO(U, MAXT, MINB) \
  return launch_gn_v<T, U, HB, MAXT, MINB>(g, x, pre_bias, gamma, beta, y, ws, N, HW, C, groups, eps, silu, stream)
  if (variant == 1) {
    if (small) PAID_GN_GO(4, 256, 4);
    PAID_GN_GO(4, 512, 2);
  }
  if (small) PAID_GN_GO(8, 256, 2);
  PAID_GN_GO(8, 512, 1);
#undef PAID_GN_GO
}

template <typename T>
int launch_gn_t(const void* x, const void* pre_bias, const void* gamma, const void* beta, void* y, float* ws, int N,
                long long HW, int C, int groups, float eps, int silu, cudaStream_t stream) {
  return pre_bias ? launch_gn_hb<T, true>(x, pre_bias, gamma, beta, y, ws, N, HW, C, groups, eps, silu, stream)
                  : launch_gn_hb<T, false>(x, pre_bias, gamma, beta, y, ws, N, HW, C, groups, eps, silu, stream);
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

template <typename T, int NV>
static void launch_ln_nv(const void* x, const void* delta, const void* gamma, const void* beta, void* x_out, void* h_out,
                         long long rows, int C, float eps, cudaStream_t stream) {
  const unsigned blocks = (unsigned)((rows + kLnWarps - 1) / kLnWarps);
  add_layer_norm_kernel<T, NV><<<blocks, kLnWarps * 32, 0, stream>>>((const T*)x, (const T*)delta, (const T*)gamma, (const T*)beta,
                                                                   (T*)x_out, (T*)h_out, rows, C, eps);
}

template <typename T>
static void launch_ln_t(const void* x, const void* delta, const void* gamma, const void* beta, void* x_out, void* h_out,
                        long long rows, int C, float eps, cudaStream_t stream) {
  const int nv = (C / 8 + 31) / 32;   // vectors per lane
  if (nv <= 1) launch_ln_nv<T, 1>(x, delta, gamma, beta, x_out, h_out, rows, C, eps, stream);
  else if (nv <= 2) launch_ln_nv<T, 2>(x, delta, gamma, beta, x_out, h_out, rows, C, eps, stream);
  else if (nv <= 3) launch_ln_nv<T, 3>(x, delta, gamma, beta, x_out, h_out, rows, C, eps, stream);
  else if (nv <= 5) launch_ln_nv<T, 5>(x, delta, gamma, beta, x_out, h_out, rows, C, eps, stream);
  else launch_ln_nv<T, kLnMaxVec>(x, delta, gamma, beta, x_out, h_out, rows, C, eps, stream);
}

int launch_add_layer_norm(const void* x, const void* delta, const void* gamma, const void* beta, void* x_out, void* h_out,
                          long long rows, int C, float eps, int dtype, cudaStream_t stream) {
  if (dtype == PAID_F16) launch_ln_t<__half>(x, delta, gamma, beta, x_out, h_out, rows, C, eps, stream);
  else launch_ln_t<__nv_bfloat16>(x, delta, gamma, beta, x_out, h_out, rows, C, eps, stream);
  PAID_LAUNCH_CHECK("add_layer_norm_kernel");
  return PAID_OK;
}

int launch_residual_bias_add(const void* a, const void* b, const void* bias, void* out, long long rows, int C, int dtype,
                             cudaStream_t stream) {
  const long long total = rows * (C / 8);
  if (total >= (1LL << 31) - (148LL * 8 * 256 * 4)) return fail(PAID_EUNSUPPORTED, "paid_residual_bias_add: rows * C / 8 must be below 2^31");
  long long blocks = (total + 256 * 4 - 1) / (256 * 4);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  if (dtype == PAID_F16)
    residual_bias_add_kernel<__half><<<(unsigned)blocks, 256, 0, stream>>>((const __half*)a, (const __half*)b, (const __half*)bias,
                                                                          (__half*)out, (unsigned)total, (unsigned)(C / 8));
  else
    residual_bias_add_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, stream>>>(
        (const __nv_bfloat16*)a, (const __nv_bfloat16*)b, (const __nv_bfloat16*)bias, (__nv_bfloat16*)out, (unsigned)total,
        (unsigned)(C / 8));
  PAID_LAUNCH_CHECK("residual_bias_add_kernel");
  return PAID_OK;
}

}  // namespace paid
