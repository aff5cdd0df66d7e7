// Projection GEMM on the 5th-generation tensor cores:  y (M,N) = x (M,K) w(N,K)^T + bias.
//
// Replaces attn.to_q / to_k / to_v / to_out[0] (reference interpolation.py:613, 623-624, 666).
// Both operands are K-major, so x and w tiles are TMA-loaded as they lie in HBM (128-byte swizzle),
// multiplied with tcgen05.mma (128 x BN x 16 per instruction, fp32 accumulator in TMEM) and written back
// by four epilogue warps straight from TMEM (+bias, -> fp16/bf16).
//
// Persistent, warp-specialised: one CTA per SM walks the output tiles (n fastest, so concurrently running
// CTAs share the same rows of x in L2).  warp 0 = TMA producer (4-stage ring), warp 1 = MMA issuer,
// warps 2-9 = epilogue (two per TMEM lane quadrant, half of the columns each).  The TMEM accumulator is DOUBLE-BUFFERED (2 x BN columns): the epilogue of tile i
// overlaps the main loop of tile i+1.  BN = 256 (SS-mode operand fetch per flop is half of a 128-wide tile:
// the 1-CTA MMA is shared-memory-bandwidth bound) unless the output width is not a multiple of 256.
// Up to three weight matrices that share the same input (q/k/v of a self-attention layer, k/v of a
// cross-attention layer) run as ONE launch: the tile index also enumerates the group.
#include <cstdlib>
#include <type_traits>

#include "paid_common.cuh"
#include "sm100_ptx.cuh"

namespace paid {
namespace {

constexpr int BM = 128, BK = 64;
// warp 0: TMA producer, warp 1: MMA issuer, warps 2-9: EIGHT epilogue warps -- two per TMEM lane quadrant, each draining half
// of the accumulator columns.  With four, the GEGLU epilogue (two TMEM loads, 32 erf-GELUs and a store per 32 columns) took
// longer than the K = 640 main loop of the 64x64-level feed-forward (tensor pipe 39 % busy, profiles/r2_ncu_gemm.txt).
constexpr int kGemmThreads = 320;
constexpr int A_BYTES = BM * BK * 2;

struct GroupPtrs {
  const void* bias[3];
  void* y[3];
  int wide;   // every y is 32-byte aligned and N % 16 == 0: rows are written with 32-byte stores (whole DRAM sectors)
};

// 16 consecutive 16-bit outputs (8 packed words) of one row: one 32-byte store, or two 16-byte stores
__device__ __forceinline__ void store16(void* dst, const uint32_t (&o)[8], bool wide) {
  if (wide) {
    ptx::st_global_256(dst, o);
  } else {
    reinterpret_cast<uint4*>(dst)[0] = make_uint4(o[0], o[1], o[2], o[3]);
    reinterpret_cast<uint4*>(dst)[1] = make_uint4(o[4], o[5], o[6], o[7]);
  }
}

// 32 accumulator columns of one row -> y (+ bias); columns >= N (ragged last tile, N % 8 == 0) are not written
template <typename T>
__device__ __forceinline__ void store_row32(T* __restrict__ dst, const uint32_t (&r)[32], const T* __restrict__ bias, int col0,
                                            int N, bool wide) {
#pragma unroll
  for (int v = 0; v < 2; ++v) {  // 16 columns per step
    const int col = col0 + v * 16;
    if (col >= N) break;
    uint32_t o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float f0 = __uint_as_float(r[v * 16 + 2 * j]), f1 = __uint_as_float(r[v * 16 + 2 * j + 1]);
      if (bias) { f0 += to_f32(bias[col + 2 * j]); f1 += to_f32(bias[col + 2 * j + 1]); }
      o[j] = pack2<T>(f0, f1);
    }
    if (col + 16 <= N) store16(dst + v * 16, o, wide);
    else *reinterpret_cast<uint4*>(dst + v * 16) = make_uint4(o[0], o[1], o[2], o[3]);   // 8 columns left
  }
}

// 32 output columns of a GEGLU tile: out = (a + ba) * gelu(g + bg); bias is (2 N,) = [ba ; bg] or NULL
template <typename T>
__device__ __forceinline__ void geglu_store(T* __restrict__ dst, const uint32_t (&ra)[32], const uint32_t (&rg)[32],
                                            const T* __restrict__ bias, int col0, int N, bool wide) {
#pragma unroll
  for (int v = 0; v < 2; ++v) {  // 16 columns per step
    const int col = col0 + v * 16;
    if (col >= N) break;
    uint32_t o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float a0 = __uint_as_float(ra[v * 16 + 2 * j]), a1 = __uint_as_float(ra[v * 16 + 2 * j + 1]);
      float g0 = __uint_as_float(rg[v * 16 + 2 * j]), g1 = __uint_as_float(rg[v * 16 + 2 * j + 1]);
      if (bias) {
        a0 += to_f32(bias[col + 2 * j]); a1 += to_f32(bias[col + 2 * j + 1]);
        g0 += to_f32(bias[N + col + 2 * j]); g1 += to_f32(bias[N + col + 2 * j + 1]);
      }
      o[j] = pack2<T>(a0 * gelu_erf(g0), a1 * gelu_erf(g1));
    }
    if (col + 16 <= N) store16(dst + v * 16, o, wide);
    else *reinterpret_cast<uint4*>(dst + v * 16) = make_uint4(o[0], o[1], o[2], o[3]);   // 8 columns left
  }
}

template <int BN> struct Cfg {
  static constexpr int kStages = BN == 256 ? 4 : 6;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = A_BYTES + kBBytes;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /* alignment slack */ + 256 /* barriers */;
  static constexpr int kTmemCols = 2 * BN;  // two accumulators
};

// GEGLU = true (BN = 256 only): w is (2 N, K) = [Wa ; Wg]; an output tile is 128 columns wide, its weight tile is rows
// [n0, n0 + 128) of Wa and of Wg, so accumulator columns 0..127 hold a and 128..255 hold g of the SAME output columns and
// the epilogue writes (a + ba) * gelu(g + bg): the feed-forward's first Linear and its GEGLU in one kernel.
template <typename T, int BN, bool GEGLU>
__global__ void __launch_bounds__(kGemmThreads, 1)
linear_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB0,
                 const __grid_constant__ CUtensorMap tmB1, const __grid_constant__ CUtensorMap tmB2,
                 const GroupPtrs gp, long long M, int N, int K, int m_tiles, int n_tiles, int total_tiles) {
  static_assert(!GEGLU || BN == 256, "the GEGLU epilogue pairs two 128-column halves");
  constexpr int TILE_N = GEGLU ? 128 : BN;   // output columns per tile
  using C = Cfg<BN>;
  constexpr int ST = C::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + ST * C::kStageBytes);
  uint64_t* empty = full + ST;
  uint64_t* acc_full = empty + ST;      // [2] MMA -> epilogue
  uint64_t* acc_empty = acc_full + 2;   // [2] epilogue -> MMA (4 warps arrive)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_kb = (K + BK - 1) / BK;

  if (warp == 0 && ptx::elect_one()) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB0);
    for (int s = 0; s < ST; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { ptx::mbar_init(&acc_full[b], 1); ptx::mbar_init(&acc_empty[b], 8); }
    ptx::fence_barrier_init();
  }
  if (warp == 1) { ptx::tmem_alloc(tmem_slot, C::kTmemCols); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  ptx::pdl_launch_dependents();  // the next kernel may begin its prologue
  ptx::pdl_wait();               // ... and this one may not read its inputs before its predecessor is done

  // tile -> (group, m block, n block); n fastest
  auto decode = [&](int tile, int& group, long long& m0, int& n0) {
    n0 = (tile % n_tiles) * TILE_N;
    const int r = tile / n_tiles;
    m0 = (long long)(r % m_tiles) * BM;
    group = r / m_tiles;
  };

  if (warp == 0) {
    if (ptx::elect_one()) {
      int kc = 0;  // k-blocks loaded so far (ring position)
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int group, n0; long long m0;
        decode(tile, group, m0, n0);
        const CUtensorMap* tmB = group == 0 ? &tmB0 : (group == 1 ? &tmB1 : &tmB2);
        for (int kb = 0; kb < num_kb; ++kb, ++kc) {
          const int s = kc % ST;
          ptx::mbar_wait(&empty[s], ((kc / ST) & 1) ^ 1);
          ptx::mbar_arrive_expect_tx(&full[s], C::kStageBytes);
          uint8_t* a = smem + s * C::kStageBytes;
          ptx::tma_load_2d(a, &tmA, &full[s], kb * BK, (int)m0);
          ptx::tma_load_2d(a + A_BYTES, tmB, &full[s], kb * BK, n0);
          if (BN == 256) ptx::tma_load_2d(a + A_BYTES + 128 * BK * 2, tmB, &full[s], kb * BK, GEGLU ? N + n0 : n0 + 128);
        }
      }
    }
  } else if (warp == 1) {
    if (ptx::elect_one()) {
      constexpr uint32_t idesc = ptx::make_idesc(BM, BN, std::is_same<T, __nv_bfloat16>::value ? 1 : 0, 0);
      int kc = 0, tc = 0;  // k-blocks / tiles issued so far
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tc) {
        const int ab = tc & 1;  // accumulator buffer
        ptx::mbar_wait(&acc_empty[ab], ((tc >> 1) & 1) ^ 1);  // epilogue has drained this accumulator
        ptx::tc_fence_after();
        for (int kb = 0; kb < num_kb; ++kb, ++kc) {
          const int s = kc % ST;
          ptx::mbar_wait(&full[s], (kc / ST) & 1);
          ptx::tc_fence_after();
          const uint32_t a = ptx::smem_u32(smem + s * C::kStageBytes);
          const uint64_t adesc = ptx::make_smem_desc_sw128(a, 16, 1024);
          const uint64_t bdesc = ptx::make_smem_desc_sw128(a + A_BYTES, 16, 1024);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)  // +32 bytes per K step inside the 128-byte swizzle atom
            ptx::mma_ss(tmem + ab * BN, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          ptx::tc_commit(&empty[s]);
        }
        ptx::tc_commit(&acc_full[ab]);
      }
    }
  } else {
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    const int half = (warp - 2) >> 2;   // which half of the accumulator columns this warp drains
    int tc = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tc) {
      int group, n0; long long m0;
      decode(tile, group, m0, n0);
      const T* __restrict__ bias = (const T*)gp.bias[group];
      T* __restrict__ y = (T*)gp.y[group];
      const int ab = tc & 1;
      ptx::mbar_wait(&acc_full[ab], (tc >> 1) & 1);
      ptx::tc_fence_after();
      const long long row = m0 + quad * 32 + lane;
      if constexpr (GEGLU) {
#pragma unroll 1
        for (int c = half * 2; c < half * 2 + 2; ++c) {
          uint32_t ra[32], rg[32];
          ptx::tmem_ld32(tmem + ((uint32_t)(quad * 32) << 16) + ab * BN + c * 32, ra);
          ptx::tmem_ld32(tmem + ((uint32_t)(quad * 32) << 16) + ab * BN + 128 + c * 32, rg);
          ptx::tmem_wait_ld();
          const int col0 = n0 + c * 32;
          if (row < M && col0 < N) geglu_store<T>(y + row * N + col0, ra, rg, bias, col0, N, gp.wide != 0);
        }
      } else {
#pragma unroll 1
      for (int c = half * (BN / 64); c < (half + 1) * (BN / 64); ++c) {
        uint32_t r[32];
        ptx::tmem_ld32(tmem + ((uint32_t)(quad * 32) << 16) + ab * BN + c * 32, r);
        ptx::tmem_wait_ld();
        const int col0 = n0 + c * 32;
        if (row < M && col0 < N) {
          store_row32<T>(y + row * N + col0, r, bias, col0, N, gp.wide != 0);
        }
      }
      }
      // this accumulator buffer may be overwritten by the tile after next
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&acc_empty[ab]);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) { __syncwarp(); ptx::tmem_dealloc(tmem, C::kTmemCols); }
}

template <typename T, int BN, bool GEGLU = false>
int launch_t(const CUtensorMap& tmA, const CUtensorMap* tmB, const GroupPtrs& gp, int groups, long long M, int N, int K,
             cudaStream_t stream) {
  auto kern = linear_tc_kernel<T, BN, GEGLU>;
  int num_sms = 0;
  PAID_CUDA_CHECK(ensure_kernel_configured((const void*)kern, Cfg<BN>::kSmemBytes, &num_sms));
  constexpr int TILE_N = GEGLU ? 128 : BN;
  const int m_tiles = (int)((M + BM - 1) / BM), n_tiles = (N + TILE_N - 1) / TILE_N;
  const int total = m_tiles * n_tiles * groups;
  dim3 grid(total < num_sms ? total : num_sms);
  PAID_CUDA_CHECK(launch_pdl(kern, grid, dim3(kGemmThreads), Cfg<BN>::kSmemBytes, stream, tmA, tmB[0], tmB[1], tmB[2], gp, M, N, K,
                             m_tiles, n_tiles, total));
  PAID_LAUNCH_CHECK("linear_tc_kernel");
  return PAID_OK;
}

// ------------------------------------------------------------------------------------------------------
// CTA-pair variant (tcgen05 cta_group::2): a cluster of two CTAs on one TPC computes a 256 x 256 tile.  Each
// CTA loads 128 rows of x and HALF (128 rows) of the w tile, the leader CTA issues 256x256x16 MMAs that read
// the operands of both CTAs, each CTA's TMEM receives its own 128 accumulator rows.  Per SM the operand
// traffic through shared memory is halved relative to a 1-CTA 128x256 tile, which is what bounds that one.
// ------------------------------------------------------------------------------------------------------
constexpr int P_STAGES = 6;
constexpr int P_STAGE_BYTES = A_BYTES + 128 * BK * 2;  // per CTA: 128 rows of x + 128 rows of w
constexpr int P_SMEM_BYTES = P_STAGES * P_STAGE_BYTES + 1024 + 256;

template <typename T, bool GEGLU>
__global__ void __launch_bounds__(kGemmThreads, 1)
linear_tc_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB0,
                      const __grid_constant__ CUtensorMap tmB1, const __grid_constant__ CUtensorMap tmB2,
                      const GroupPtrs gp, long long M, int N, int K, int m_tiles, int n_tiles, int total_tiles) {
  constexpr int ST = P_STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + ST * P_STAGE_BYTES);  // used in the leader CTA
  uint64_t* empty = full + ST;                                              // per CTA (multicast commit)
  uint64_t* acc_full = empty + ST;                                          // per CTA (multicast commit)
  uint64_t* acc_empty = acc_full + 2;                                       // used in the leader CTA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int num_kb = (K + BK - 1) / BK;

  if (warp == 0 && ptx::elect_one()) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB0);
    for (int s = 0; s < ST; ++s) { ptx::mbar_init(&full[s], 2); ptx::mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { ptx::mbar_init(&acc_full[b], 1); ptx::mbar_init(&acc_empty[b], 16); }
    ptx::fence_barrier_init();
  }
  if (warp == 1) { ptx::tmem_alloc_pair(tmem_slot, 512); ptx::tmem_relinquish_pair(); }
  ptx::tc_fence_before();
  ptx::cluster_sync();  // barriers and TMEM of both CTAs are ready
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();

  // GEGLU: an output tile is 256 rows x 128 columns; CTA 0 holds rows [n0, n0 + 128) of Wa, CTA 1 the same rows of Wg
  // (w is (2 N, K) = [Wa ; Wg]), so accumulator columns 0..127 are a and 128..255 are g of the same output columns
  constexpr int TILE_N = GEGLU ? 128 : 256;
  auto decode = [&](int tile, int& group, long long& m0, int& n0) {
    n0 = (tile % n_tiles) * TILE_N;
    const int r = tile / n_tiles;
    m0 = (long long)(r % m_tiles) * 256;
    group = r / m_tiles;
  };

  if (warp == 0) {
    if (ptx::elect_one()) {
      int kc = 0;
      for (int tile = pair; tile < total_tiles; tile += npairs) {
        int group, n0; long long m0;
        decode(tile, group, m0, n0);
        const CUtensorMap* tmB = group == 0 ? &tmB0 : (group == 1 ? &tmB1 : &tmB2);
        for (int kb = 0; kb < num_kb; ++kb, ++kc) {
          const int s = kc % ST;
          ptx::mbar_wait(&empty[s], ((kc / ST) & 1) ^ 1);  // this CTA's slot has been consumed
          if (rank == 0) ptx::mbar_arrive_expect_tx(&full[s], 2 * P_STAGE_BYTES);  // bytes of both CTAs
          else ptx::mbar_arrive_cluster(ptx::mapa(ptx::smem_u32(&full[s]), 0));
          uint8_t* a = smem + s * P_STAGE_BYTES;
          ptx::tma_load_2d_pair(a, &tmA, &full[s], kb * BK, (int)m0 + (int)rank * 128);
          ptx::tma_load_2d_pair(a + A_BYTES, tmB, &full[s], kb * BK, GEGLU ? n0 + (int)rank * N : n0 + (int)rank * 128);
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0 && ptx::elect_one()) {
      constexpr uint32_t idesc = ptx::make_idesc(256, 256, std::is_same<T, __nv_bfloat16>::value ? 1 : 0, 0);
      int kc = 0, tc = 0;
      for (int tile = pair; tile < total_tiles; tile += npairs, ++tc) {
        const int ab = tc & 1;
        ptx::mbar_wait(&acc_empty[ab], ((tc >> 1) & 1) ^ 1);  // the epilogues of both CTAs have drained it
        ptx::tc_fence_after();
        for (int kb = 0; kb < num_kb; ++kb, ++kc) {
          const int s = kc % ST;
          ptx::mbar_wait(&full[s], (kc / ST) & 1);
          ptx::tc_fence_after();
          const uint32_t a = ptx::smem_u32(smem + s * P_STAGE_BYTES);
          const uint64_t adesc = ptx::make_smem_desc_sw128(a, 16, 1024);
          const uint64_t bdesc = ptx::make_smem_desc_sw128(a + A_BYTES, 16, 1024);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            ptx::mma_ss_pair(tmem + ab * 256, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          ptx::tc_commit_pair(&empty[s]);   // frees the slot in both CTAs
        }
        ptx::tc_commit_pair(&acc_full[ab]);  // both CTAs' epilogues may read their rows
      }
    }
  } else {
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;   // which half of the accumulator columns this warp drains
    int tc = 0;
    for (int tile = pair; tile < total_tiles; tile += npairs, ++tc) {
      int group, n0; long long m0;
      decode(tile, group, m0, n0);
      const T* __restrict__ bias = (const T*)gp.bias[group];
      T* __restrict__ y = (T*)gp.y[group];
      const int ab = tc & 1;
      ptx::mbar_wait(&acc_full[ab], (tc >> 1) & 1);
      ptx::tc_fence_after();
      const long long row = m0 + rank * 128 + quad * 32 + lane;
      if constexpr (GEGLU) {
#pragma unroll 1
        for (int c = half * 2; c < half * 2 + 2; ++c) {
          uint32_t ra[32], rg[32];
          ptx::tmem_ld32(tmem + ((uint32_t)(quad * 32) << 16) + ab * 256 + c * 32, ra);
          ptx::tmem_ld32(tmem + ((uint32_t)(quad * 32) << 16) + ab * 256 + 128 + c * 32, rg);
          ptx::tmem_wait_ld();
          const int col0 = n0 + c * 32;
          if (row < M && col0 < N) geglu_store<T>(y + row * N + col0, ra, rg, bias, col0, N, gp.wide != 0);
        }
      } else {
#pragma unroll 1
      for (int c = half * 4; c < half * 4 + 4; ++c) {
        uint32_t r[32];
        ptx::tmem_ld32(tmem + ((uint32_t)(quad * 32) << 16) + ab * 256 + c * 32, r);
        ptx::tmem_wait_ld();
        const int col0 = n0 + c * 32;
        if (row < M && col0 < N) {
          store_row32<T>(y + row * N + col0, r, bias, col0, N, gp.wide != 0);
        }
      }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_cluster(ptx::mapa(ptx::smem_u32(&acc_empty[ab]), 0));
    }
  }
  ptx::tc_fence_before();
  ptx::cluster_sync();  // neither CTA may exit (or free TMEM) while the other can still touch it
  if (warp == 1) { __syncwarp(); ptx::tmem_dealloc_pair(tmem, 512); }
}

template <typename T, bool GEGLU = false>
int launch_pair_t(const CUtensorMap& tmA, const CUtensorMap* tmB, const GroupPtrs& gp, int groups, long long M, int N,
                  int K, cudaStream_t stream) {
  auto kern = linear_tc_pair_kernel<T, GEGLU>;
  int num_sms = 0;
  PAID_CUDA_CHECK(ensure_kernel_configured((const void*)kern, P_SMEM_BYTES, &num_sms));
  const int m_tiles = (int)((M + 255) / 256), n_tiles = GEGLU ? (N + 127) / 128 : (N + 255) / 256;
  const int total = m_tiles * n_tiles * groups;
  int pairs = num_sms / 2;
  if (total < pairs) pairs = total;
  PAID_CUDA_CHECK(launch_pdl_pairs(kern, dim3(2 * pairs), dim3(kGemmThreads), P_SMEM_BYTES, stream, tmA, tmB[0], tmB[1], tmB[2], gp,
                                   M, N, K, m_tiles, n_tiles, total));
  PAID_LAUNCH_CHECK("linear_tc_pair_kernel");
  return PAID_OK;
}

}  // namespace

bool linear_tc_supported(long long M, int Nout, int K) {
  return M >= 1 && Nout % 8 == 0 && K % 8 == 0 && (M + BM - 1) / BM < (1 << 22);
}

int launch_linear_tc_grouped(const void* x, const void* const* w, const void* const* bias, void* const* y, int groups,
                             long long M, int Nout, int K, int dtype, cudaStream_t stream) {
  if (groups < 1 || groups > 3) return fail(PAID_EINVAL, "linear: 1..3 weight groups per launch");
  CUtensorMap tmA, tmB[3];
  GroupPtrs gp{};
  if ((uintptr_t)x & 15) return fail(PAID_EINVAL, "linear: pointers must be 16-byte aligned");
  int st = make_tmap_2d(&tmA, x, dtype, M, K, K, BM);
  if (st != PAID_OK) return st;
  for (int g = 0; g < 3; ++g) {
    const int s = g < groups ? g : 0;  // unused slots repeat group 0 (any valid descriptor)
    if (((uintptr_t)w[s] | (uintptr_t)y[s]) & 15) return fail(PAID_EINVAL, "linear: pointers must be 16-byte aligned");
    if ((st = make_tmap_2d(&tmB[g], w[s], dtype, Nout, K, K, 128)) != PAID_OK) return st;  // 128-row boxes
    gp.bias[g] = bias ? bias[s] : nullptr;
    gp.y[g] = y[s];
  }
  gp.wide = Nout % 16 == 0 && !(((uintptr_t)gp.y[0] | (uintptr_t)gp.y[1] | (uintptr_t)gp.y[2]) & 31);
  const bool wide = Nout % 256 == 0;
  // CTA pairs whenever the 256-wide tiles are at least 5/6 full (640 = 2.5 tiles still beats the 1-CTA kernel)
  const bool pairs_ok = M >= 256 && Nout >= 256 && 6 * Nout >= 5 * 256 * ((Nout + 255) / 256);
  static const bool no_pairs = getenv("PAID_NO_CTA_PAIRS") != nullptr;   // debugging knob, read once
  if (pairs_ok && !no_pairs)
    return dtype == PAID_F16 ? launch_pair_t<__half>(tmA, tmB, gp, groups, M, Nout, K, stream)
                             : launch_pair_t<__nv_bfloat16>(tmA, tmB, gp, groups, M, Nout, K, stream);
  if (dtype == PAID_F16)
    return wide ? launch_t<__half, 256>(tmA, tmB, gp, groups, M, Nout, K, stream)
                : launch_t<__half, 128>(tmA, tmB, gp, groups, M, Nout, K, stream);
  return wide ? launch_t<__nv_bfloat16, 256>(tmA, tmB, gp, groups, M, Nout, K, stream)
              : launch_t<__nv_bfloat16, 128>(tmA, tmB, gp, groups, M, Nout, K, stream);
}

bool linear_geglu_tc_supported(long long M, int D, int K) {
  return M >= 1 && D % 8 == 0 && K % 8 == 0 && (M + BM - 1) / BM < (1 << 22);
}

// y (M, D) = (x Wa^T + ba) * gelu(x Wg^T + bg), w = [Wa ; Wg] (2 D, K), bias = [ba ; bg] (2 D,) or NULL
int launch_linear_geglu_tc(const void* x, const void* w, const void* bias, void* y, long long M, int D, int K, int dtype,
                           cudaStream_t stream) {
  if (((uintptr_t)x | (uintptr_t)w | (uintptr_t)y | (uintptr_t)bias) & 15) return fail(PAID_EINVAL, "linear_geglu: pointers must be 16-byte aligned");
  CUtensorMap tmA, tmB[3];
  GroupPtrs gp{};
  int st = make_tmap_2d(&tmA, x, dtype, M, K, K, BM);
  if (st != PAID_OK) return st;
  if ((st = make_tmap_2d(&tmB[0], w, dtype, 2LL * D, K, K, 128)) != PAID_OK) return st;
  tmB[1] = tmB[2] = tmB[0];
  gp.bias[0] = bias; gp.y[0] = y;
  gp.wide = D % 16 == 0 && !((uintptr_t)y & 31);
  static const bool no_pairs = getenv("PAID_NO_CTA_PAIRS") != nullptr;
  if (M >= 256 && D >= 128 && !no_pairs)
    return dtype == PAID_F16 ? launch_pair_t<__half, true>(tmA, tmB, gp, 1, M, D, K, stream)
                             : launch_pair_t<__nv_bfloat16, true>(tmA, tmB, gp, 1, M, D, K, stream);
  return dtype == PAID_F16 ? launch_t<__half, 256, true>(tmA, tmB, gp, 1, M, D, K, stream)
                           : launch_t<__nv_bfloat16, 256, true>(tmA, tmB, gp, 1, M, D, K, stream);
}

int launch_linear_tc(const void* x, const void* w, const void* bias, void* y, long long M, int Nout, int K, int dtype,
                     cudaStream_t stream) {
  return launch_linear_tc_grouped(x, &w, bias ? &bias : nullptr, &y, 1, M, Nout, K, dtype, stream);
}

}  // namespace paid
