// Generic-shape CUDA kernels: any head_dim <= 160, any L, any K.  They serve the geometries the
// tcgen05 kernels do not cover yet (SD1.5 head dims 40/80/160) and are the on-device cross-check
// of the tcgen05 path (PAID_FLAG_GENERIC_KERNELS).  Plain SIMT fp32 math, no tensor cores.
#include "paid_common.cuh"

namespace paid {

// ------------------------------------------------------------------------------------------------
// y (M,Nout) = x (M,K) w(Nout,K)^T + bias        64x64 tile, 16-deep k slices, 4x4 outputs per thread
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) linear_generic_kernel(const T* __restrict__ x, const T* __restrict__ w,
                                                             const T* __restrict__ bias, T* __restrict__ y,
                                                             long long M, int Nout, int K) {
  __shared__ float As[16][64 + 4];
  __shared__ float Bs[16][64 + 4];
  const int tid = threadIdx.x;
  const long long m0 = (long long)blockIdx.y * 64;
  const int n0 = blockIdx.x * 64;
  const int tr = tid / 16, tc = tid % 16;  // 16x16 threads, each a 4x4 patch
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += 16) {
    // 64 rows x 16 k = 1024 elements per operand, 4 per thread
    for (int i = tid; i < 1024; i += 256) {
      int r = i / 16, kk = i % 16;
      long long gm = m0 + r;
      int gn = n0 + r, gk = k0 + kk;
      As[kk][r] = (gm < M && gk < K) ? to_f32(x[gm * K + gk]) : 0.f;
      Bs[kk][r] = (gn < Nout && gk < K) ? to_f32(w[(long long)gn * K + gk]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[kk][tr * 4 + i]; b[i] = Bs[kk][tc * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    long long gm = m0 + tr * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int gn = n0 + tc * 4 + j;
      if (gn < Nout) y[gm * Nout + gn] = from_f32<T>(acc[i][j] + (bias ? to_f32(bias[gn]) : 0.f));
    }
  }
}

int launch_linear_generic(const void* x, const void* w, const void* bias, void* y, long long M, int Nout, int K,
                          int dtype, cudaStream_t stream) {
  dim3 grid((Nout + 63) / 64, (unsigned)((M + 63) / 64));
  if (dtype == PAID_F16)
    linear_generic_kernel<__half><<<grid, 256, 0, stream>>>((const __half*)x, (const __half*)w, (const __half*)bias,
                                                            (__half*)y, M, Nout, K);
  else
    linear_generic_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(
        (const __nv_bfloat16*)x, (const __nv_bfloat16*)w, (const __nv_bfloat16*)bias, (__nv_bfloat16*)y, M, Nout, K);
  PAID_LAUNCH_CHECK("linear_generic_kernel");
  return PAID_OK;
}

// ------------------------------------------------------------------------------------------------
// y (M,D) = (x Wa^T + ba) * gelu(x Wg^T + bg), w = [Wa ; Wg] (2D,K): any-shape counterpart of the GEGLU epilogue of
// gemm_tc.cu (cross-check, and K / D that TMA cannot address)
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) linear_geglu_generic_kernel(const T* __restrict__ x, const T* __restrict__ w,
                                                                   const T* __restrict__ bias, T* __restrict__ y,
                                                                   long long M, int D, int K) {
  __shared__ float As[16][64 + 4];
  __shared__ float Ba[16][64 + 4];
  __shared__ float Bg[16][64 + 4];
  const int tid = threadIdx.x;
  const long long m0 = (long long)blockIdx.y * 64;
  const int n0 = blockIdx.x * 64;
  const int tr = tid / 16, tc = tid % 16;
  float acc_a[4][4] = {}, acc_g[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += 16) {
    for (int i = tid; i < 1024; i += 256) {
      int r = i / 16, kk = i % 16;
      long long gm = m0 + r;
      int gn = n0 + r, gk = k0 + kk;
      As[kk][r] = (gm < M && gk < K) ? to_f32(x[gm * K + gk]) : 0.f;
      Ba[kk][r] = (gn < D && gk < K) ? to_f32(w[(long long)gn * K + gk]) : 0.f;
      Bg[kk][r] = (gn < D && gk < K) ? to_f32(w[(long long)(D + gn) * K + gk]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a[4], ba[4], bg[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[kk][tr * 4 + i]; ba[i] = Ba[kk][tc * 4 + i]; bg[i] = Bg[kk][tc * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc_a[i][j] = fmaf(a[i], ba[j], acc_a[i][j]);
          acc_g[i][j] = fmaf(a[i], bg[j], acc_g[i][j]);
        }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    long long gm = m0 + tr * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int gn = n0 + tc * 4 + j;
      if (gn < D) {
        const float av = acc_a[i][j] + (bias ? to_f32(bias[gn]) : 0.f);
        const float gv = acc_g[i][j] + (bias ? to_f32(bias[D + gn]) : 0.f);
        y[gm * D + gn] = from_f32<T>(av * gelu_erf(gv));
      }
    }
  }
}

int launch_linear_geglu_generic(const void* x, const void* w, const void* bias, void* y, long long M, int D, int K, int dtype,
                                cudaStream_t stream) {
  dim3 grid((D + 63) / 64, (unsigned)((M + 63) / 64));
  if (dtype == PAID_F16)
    linear_geglu_generic_kernel<__half><<<grid, 256, 0, stream>>>((const __half*)x, (const __half*)w, (const __half*)bias,
                                                                  (__half*)y, M, D, K);
  else
    linear_geglu_generic_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(
        (const __nv_bfloat16*)x, (const __nv_bfloat16*)w, (const __nv_bfloat16*)bias, (__nv_bfloat16*)y, M, D, K);
  PAID_LAUNCH_CHECK("linear_geglu_generic_kernel");
  return PAID_OK;
}

// ------------------------------------------------------------------------------------------------
// endpoint lerp for INNER mode: kx[n] = (1-c_n) kb + c_n ke, vx likewise  (interpolation.py:772-775)
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void lerp_endpoints_kernel(const T* __restrict__ kb, const T* __restrict__ vb, const T* __restrict__ ke,
                                      const T* __restrict__ ve, const float* __restrict__ coef, T* __restrict__ kx,
                                      T* __restrict__ vx, long long LC) {
  const int n = blockIdx.y;
  const float c = coef[n];
  const long long stride = (long long)gridDim.x * blockDim.x * 2;
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 2; i < LC; i += stride) {
    // LC is even (C is a multiple of 8): two elements per thread per step
    float k0 = to_f32(kb[i]), k1 = to_f32(kb[i + 1]), e0 = to_f32(ke[i]), e1 = to_f32(ke[i + 1]);
    float v0 = to_f32(vb[i]), v1 = to_f32(vb[i + 1]), f0 = to_f32(ve[i]), f1 = to_f32(ve[i + 1]);
    uint32_t pk = pack2<T>((1.f - c) * k0 + c * e0, (1.f - c) * k1 + c * e1);
    uint32_t pv = pack2<T>((1.f - c) * v0 + c * f0, (1.f - c) * v1 + c * f1);
    *reinterpret_cast<uint32_t*>(kx + (long long)n * LC + i) = pk;
    *reinterpret_cast<uint32_t*>(vx + (long long)n * LC + i) = pv;
  }
}

int launch_lerp_endpoints(const void* kb, const void* vb, const void* ke, const void* ve, const float* coef,
                          void* kx, void* vx, int N, long long LC, int dtype, cudaStream_t stream) {
  int bx = (int)((LC / 2 + 255) / 256);
  if (bx > 592) bx = 592;  // 4 x 148 SMs
  if (bx < 1) bx = 1;
  dim3 grid(bx, N);
  if (dtype == PAID_F16)
    lerp_endpoints_kernel<__half><<<grid, 256, 0, stream>>>((const __half*)kb, (const __half*)vb, (const __half*)ke,
                                                            (const __half*)ve, coef, (__half*)kx, (__half*)vx, LC);
  else
    lerp_endpoints_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(
        (const __nv_bfloat16*)kb, (const __nv_bfloat16*)vb, (const __nv_bfloat16*)ke, (const __nv_bfloat16*)ve, coef,
        (__nv_bfloat16*)kx, (__nv_bfloat16*)vx, LC);
  PAID_LAUNCH_CHECK("lerp_endpoints_kernel");
  return PAID_OK;
}

// ------------------------------------------------------------------------------------------------
// GEGLU: out[m, j] = h[m, j] * gelu(h[m, D + j]), exact erf GELU.  HBM-bound: 16-byte loads / stores,
// 6 bytes of traffic per output element.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) geglu_kernel(const T* __restrict__ h, T* __restrict__ out, unsigned total, unsigned vec_per_row) {
  // 32-bit index arithmetic (the launcher checks M * D / 8 < 2^31): a 64-bit division per vector costs more than the math
  const unsigned stride = gridDim.x * blockDim.x, D = vec_per_row * 8;
  for (unsigned i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < total; i0 += 2 * stride) {
    uint4 av[2], gv[2];
    unsigned m[2], j[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const unsigned i = i0 + u * stride;
      m[u] = i / vec_per_row;
      j[u] = (i - m[u] * vec_per_row) * 8;
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {   // four independent 16-byte loads in flight per thread
      if (i0 + u * stride < total) {
        av[u] = *reinterpret_cast<const uint4*>(h + (size_t)m[u] * 2 * D + j[u]);
        gv[u] = *reinterpret_cast<const uint4*>(h + (size_t)m[u] * 2 * D + D + j[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (i0 + u * stride >= total) break;
      const T* a8 = reinterpret_cast<const T*>(&av[u]);
      const T* g8 = reinterpret_cast<const T*>(&gv[u]);
      uint32_t o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float g0 = to_f32(g8[2 * e]), g1 = to_f32(g8[2 * e + 1]);
        const float r0 = to_f32(a8[2 * e]) * gelu_erf(g0);
        const float r1 = to_f32(a8[2 * e + 1]) * gelu_erf(g1);
        o[e] = pack2<T>(r0, r1);
      }
      *reinterpret_cast<uint4*>(out + (size_t)m[u] * D + j[u]) = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

int launch_geglu(const void* h, void* out, long long M, int D, int dtype, cudaStream_t stream) {
  const long long total = M * (D / 8);
  if (total >= (1LL << 31) - (148LL * 16 * 256 * 2)) return fail(PAID_EUNSUPPORTED, "paid_geglu: M * D / 8 must be below 2^31");
  long long blocks = (total + 511) / 512;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  if (dtype == PAID_F16)
    geglu_kernel<__half><<<(unsigned)blocks, 256, 0, stream>>>((const __half*)h, (__half*)out, (unsigned)total, (unsigned)(D / 8));
  else
    geglu_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, stream>>>((const __nv_bfloat16*)h, (__nv_bfloat16*)out,
                                                                      (unsigned)total, (unsigned)(D / 8));
  PAID_LAUNCH_CHECK("geglu_kernel");
  return PAID_OK;
}

// ------------------------------------------------------------------------------------------------
// generic interpolated attention.  Block = 8 warps x 4 query rows; keys in tiles of 32 (lane = key);
// output dims owned by lanes (j = lane + 32*jj).  Three slots with independent online-softmax state,
// merged at the end (paid_common.cuh).
// ------------------------------------------------------------------------------------------------
constexpr int kRowsPerWarp = 4;
constexpr int kWarps = 8;
constexpr int kRowsPerBlock = kRowsPerWarp * kWarps;  // 32
constexpr int kKeysPerTile = 32;

template <typename T, int DPL>
__global__ void __launch_bounds__(kWarps * 32) attn_generic_kernel(CoreArgs a) {
  extern __shared__ float smem[];
  const int d = a.head_dim;
  const int dp = d + 1;  // padded row pitch (floats): conflict-free lane-per-key reads
  float* Qs = smem;                          // [32][dp]  pre-scaled by scale*log2e
  float* Ks = Qs + kRowsPerBlock * dp;       // [32][dp]
  float* Vs = Ks + kKeysPerTile * dp;        // [32][dp]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = blockIdx.z, head = blockIdx.y, row0 = blockIdx.x * kRowsPerBlock;
  const int C = a.heads * d;
  const T* q = (const T*)a.q + ((long long)n * a.S) * C + head * d;
  const float qs = a.scale * kLog2e;
  for (int i = tid; i < kRowsPerBlock * d; i += blockDim.x) {
    int r = i / d, j = i % d;
    int row = row0 + r;
    Qs[r * dp + j] = row < a.S ? to_f32(q[(long long)row * C + j]) * qs : 0.f;
  }
  const float c = a.mode == PAID_PLAIN ? 0.f : a.coef[n];
  const FramePlan plan = make_frame_plan(a.mode, a.fused, n, a.begin_frame, a.end_frame, c);

  float O[3][kRowsPerWarp][DPL];
  float mS[3][kRowsPerWarp], lS[3][kRowsPerWarp];
#pragma unroll
  for (int s = 0; s < 3; ++s)
#pragma unroll
    for (int r = 0; r < kRowsPerWarp; ++r) {
      mS[s][r] = -INFINITY; lS[s][r] = 0.f;
#pragma unroll
      for (int jj = 0; jj < DPL; ++jj) O[s][r][jj] = 0.f;
    }

#pragma unroll
  for (int s = 0; s < 3; ++s) {
    const bool use = s == 0 ? plan.use0 : (s == 1 ? plan.use1 : plan.use2);
    if (!use) continue;  // block-uniform
    const T* kbase; const T* vbase;
    if (s == 0) { kbase = (const T*)a.k + n * a.stride0; vbase = (const T*)a.v + n * a.stride0; }
    else if (s == 1) { kbase = (const T*)a.k1 + n * a.stride1; vbase = (const T*)a.v1 + n * a.stride1; }
    else { kbase = (const T*)a.k2 + n * a.stride2; vbase = (const T*)a.v2 + n * a.stride2; }
    kbase += head * d; vbase += head * d;
    for (int t0 = 0; t0 < a.L; t0 += kKeysPerTile) {
      __syncthreads();
      for (int i = tid; i < kKeysPerTile * d; i += blockDim.x) {
        int r = i / d, j = i % d;
        int key = t0 + r;
        bool ok = key < a.L;
        Ks[r * dp + j] = ok ? to_f32(kbase[(long long)key * C + j]) : 0.f;
        Vs[r * dp + j] = ok ? to_f32(vbase[(long long)key * C + j]) : 0.f;
      }
      __syncthreads();
      float sc[kRowsPerWarp];
#pragma unroll
      for (int r = 0; r < kRowsPerWarp; ++r) sc[r] = 0.f;
      const float* krow = Ks + lane * dp;
      const float* qrow = Qs + (warp * kRowsPerWarp) * dp;
      for (int j = 0; j < d; ++j) {
        float kv = krow[j];
#pragma unroll
        for (int r = 0; r < kRowsPerWarp; ++r) sc[r] = fmaf(qrow[r * dp + j], kv, sc[r]);
      }
      const bool valid = t0 + lane < a.L;
#pragma unroll
      for (int r = 0; r < kRowsPerWarp; ++r) {
        float sv = valid ? sc[r] : -INFINITY;
        float tmax = sv;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
        float mnew = fmaxf(mS[s][r], tmax);
        float alpha = exp2f(mS[s][r] - mnew);  // first tile: exp2(-inf) = 0
        float p = exp2f(sv - mnew);
        float psum = p;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) psum += __shfl_xor_sync(0xffffffffu, psum, o);
        lS[s][r] = lS[s][r] * alpha + psum;
        mS[s][r] = mnew;
#pragma unroll
        for (int jj = 0; jj < DPL; ++jj) O[s][r][jj] *= alpha;
        sc[r] = p;
      }
      for (int kk = 0; kk < kKeysPerTile; ++kk) {
        float vv[DPL];
#pragma unroll
        for (int jj = 0; jj < DPL; ++jj) {
          int j = lane + 32 * jj;
          vv[jj] = j < d ? Vs[kk * dp + j] : 0.f;
        }
#pragma unroll
        for (int r = 0; r < kRowsPerWarp; ++r) {
          float pk = __shfl_sync(0xffffffffu, sc[r], kk);
#pragma unroll
          for (int jj = 0; jj < DPL; ++jj) O[s][r][jj] = fmaf(pk, vv[jj], O[s][r][jj]);
        }
      }
    }
  }

  T* out = (T*)a.out + ((long long)n * a.S) * C + head * d;
#pragma unroll
  for (int r = 0; r < kRowsPerWarp; ++r) {
    int row = row0 + warp * kRowsPerWarp + r;
    if (row >= a.S) continue;
    float cf0, cf1, cf2;
    merge_coefficients(plan, mS[0][r], lS[0][r], mS[1][r], lS[1][r], mS[2][r], lS[2][r], cf0, cf1, cf2);
    const float os = a.out_scale * (a.out_frame_scale ? a.out_frame_scale[n] : 1.f);
#pragma unroll
    for (int jj = 0; jj < DPL; ++jj) {
      int j = lane + 32 * jj;
      if (j < d) {
        float r0 = os * (cf0 * O[0][r][jj] + cf1 * O[1][r][jj] + cf2 * O[2][r][jj]);
        if (a.accumulate) r0 += to_f32(out[(long long)row * C + j]);
        out[(long long)row * C + j] = from_f32<T>(r0);
      }
    }
  }
}

template <typename T, int DPL>
static int launch_attn_generic_t(const CoreArgs& a, cudaStream_t stream) {
  size_t smem = (size_t)(kRowsPerBlock + 2 * kKeysPerTile) * (a.head_dim + 1) * sizeof(float);
  auto kern = attn_generic_kernel<T, DPL>;
  if (smem > 48 * 1024) PAID_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((a.S + kRowsPerBlock - 1) / kRowsPerBlock, a.heads, a.N);
  kern<<<grid, kWarps * 32, smem, stream>>>(a);
  PAID_LAUNCH_CHECK("attn_generic_kernel");
  return PAID_OK;
}

int launch_attn_generic(const CoreArgs& a, cudaStream_t stream) {
  if (a.head_dim < 1 || a.head_dim > 160)
    return fail(PAID_EUNSUPPORTED, "generic attention kernel supports head_dim <= 160 (got %d)", a.head_dim);
  int dpl = (a.head_dim + 31) / 32;
#define PAID_DISPATCH(DPL)                                                                                 \
  case DPL:                                                                                                \
    return a.dtype == PAID_F16 ? launch_attn_generic_t<__half, DPL>(a, stream)                             \
                               : launch_attn_generic_t<__nv_bfloat16, DPL>(a, stream);
  switch (dpl) {
    PAID_DISPATCH(1) PAID_DISPATCH(2) PAID_DISPATCH(3) PAID_DISPATCH(4) PAID_DISPATCH(5)
  }
#undef PAID_DISPATCH
  return fail(PAID_EUNSUPPORTED, "unreachable head_dim dispatch");
}

}  // namespace paid
