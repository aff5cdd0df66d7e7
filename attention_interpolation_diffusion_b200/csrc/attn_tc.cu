// Interpolated attention core on the 5th-generation tensor cores (head_dim 64).
//
// Replaces, per attention layer, the reference's endpoint replication, head split, [self ; endpoint]
// concatenations, the two materialised softmax(QK^T) matrices, the two P.V products and the alpha-lerp
// (interpolation.py:627-664 outer, 760-790 inner; deactivated 581-584) with ONE kernel launch.
//
// Work decomposition: one CTA per (frame n, head, 256 query rows) = two 128-row Q tiles.  The keys of a
// frame are up to three "slots" (own K/V, endpoint A, endpoint B; paid_common.cuh) streamed in steps of 64
// keys.  Every slot has its own fp32 output accumulator in TMEM and its own online-softmax statistics; the
// slots are merged (log-sum-exp) and alpha-lerped in the epilogue, so the endpoint K/V are read once and
// never replicated or concatenated.
//
//   warp 0        TMA producer: Q tiles once, then K / V tiles into an 8-stage shared-memory ring
//   warp 1        tcgen05.mma issuer: S_t = Q_t K^T (SS), O_{t,slot} += P_t V (A operand from TMEM)
//   warps 2-5     softmax warpgroup of Q tile 0 (one thread per query row)
//   warps 6-9     softmax warpgroup of Q tile 1
//
// TMEM (512 columns): S_0 [0,64) S_1 [64,128) (P_t aliases the low 32 columns of S_t as packed 16-bit),
// O_{t,slot} at 128 + (3 t + slot) * 64.
#include <type_traits>

#include "paid_common.cuh"
#include "sm100_ptx.cuh"

namespace paid {
namespace {

constexpr int D = 64;            // head_dim
constexpr int BM = 128;          // rows per Q tile
constexpr int QT = 2;            // Q tiles per CTA
constexpr int BN = 64;           // keys per step
constexpr int ST = 8;            // K/V ring stages
constexpr int Q_BYTES = BM * D * 2;    // 16 KB
constexpr int KV_BYTES = BN * D * 2;   // 8 KB
constexpr int SMEM_BYTES = 1024 + QT * Q_BYTES + ST * 2 * KV_BYTES + 512;
constexpr int NUM_THREADS = 32 * (2 + 4 * QT);
constexpr uint32_t TMEM_S = 0, TMEM_O = 128;
constexpr float kRescaleThreshold = 8.f;  // log2 units: rescale O only when the running max grew by > 2^8

struct TcArgs {
  int mode, fused, N, S, L, heads, begin_frame, end_frame;
  float scale_log2;
  const float* coef;
  void* out;
  int per_frame[3];  // slot K/V map has one matrix per frame (1) or a single shared matrix (0)
};

struct Barriers {
  uint64_t q_full;
  uint64_t k_full[ST], k_empty[ST], v_full[ST], v_empty[ST];
  uint64_t s_full[QT], p_full[QT];
  uint32_t tmem_slot;
};

__device__ __forceinline__ int frame_of_block(int z, int N) {
  // heavy (interior) frames first, the two cheap endpoint frames last
  if (N < 3) return z;
  return z < N - 2 ? z + 1 : (z == N - 2 ? 0 : N - 1);
}

template <typename T>
__global__ void __launch_bounds__(NUM_THREADS, 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK0,
               const __grid_constant__ CUtensorMap tmV0, const __grid_constant__ CUtensorMap tmK1,
               const __grid_constant__ CUtensorMap tmV1, const __grid_constant__ CUtensorMap tmK2,
               const __grid_constant__ CUtensorMap tmV2, const TcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                              // [QT][128][64]
  uint8_t* sK = sQ + QT * Q_BYTES;                 // [ST][64][64]
  uint8_t* sV = sK + ST * KV_BYTES;                // [ST][64][64]
  Barriers* bar = reinterpret_cast<Barriers*>(sV + ST * KV_BYTES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = frame_of_block(blockIdx.z, a.N), head = blockIdx.y, row0 = blockIdx.x * (QT * BM);
  const float c = a.mode == PAID_PLAIN ? 0.f : a.coef[n];
  const FramePlan plan = make_frame_plan(a.mode, a.fused, n, a.begin_frame, a.end_frame, c);
  const int tiles = (a.L + BN - 1) / BN;
  const int nsteps[3] = {plan.use0 ? tiles : 0, plan.use1 ? tiles : 0, plan.use2 ? tiles : 0};
  const int total_steps = nsteps[0] + nsteps[1] + nsteps[2];

  if (warp == 0 && ptx::elect_one()) {
    ptx::prefetch_tmap(&tmQ);
    ptx::prefetch_tmap(&tmK0); ptx::prefetch_tmap(&tmV0);
    ptx::mbar_init(&bar->q_full, 1);
    for (int s = 0; s < ST; ++s) {
      ptx::mbar_init(&bar->k_full[s], 1); ptx::mbar_init(&bar->k_empty[s], 1);
      ptx::mbar_init(&bar->v_full[s], 1); ptx::mbar_init(&bar->v_empty[s], 1);
    }
    for (int t = 0; t < QT; ++t) { ptx::mbar_init(&bar->s_full[t], 1); ptx::mbar_init(&bar->p_full[t], BM); }
    ptx::fence_barrier_init();
  }
  if (warp == 1) { ptx::tmem_alloc(&bar->tmem_slot, 512); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = bar->tmem_slot;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (ptx::elect_one()) {
      ptx::mbar_arrive_expect_tx(&bar->q_full, QT * Q_BYTES);
      for (int t = 0; t < QT; ++t) ptx::tma_load_4d(sQ + t * Q_BYTES, &tmQ, &bar->q_full, 0, head, row0 + t * BM, n);
      int j = 0;
      for (int slot = 0; slot < 3; ++slot) {
        const CUtensorMap* mk = slot == 0 ? &tmK0 : (slot == 1 ? &tmK1 : &tmK2);
        const CUtensorMap* mv = slot == 0 ? &tmV0 : (slot == 1 ? &tmV1 : &tmV2);
        const int fr = a.per_frame[slot] ? n : 0;
        for (int i = 0; i < nsteps[slot]; ++i, ++j) {
          const int s = j % ST;
          const uint32_t ph = (j / ST) & 1;
          ptx::mbar_wait(&bar->k_empty[s], ph ^ 1);
          ptx::mbar_arrive_expect_tx(&bar->k_full[s], KV_BYTES);
          ptx::tma_load_4d(sK + s * KV_BYTES, mk, &bar->k_full[s], 0, head, i * BN, fr);
          ptx::mbar_wait(&bar->v_empty[s], ph ^ 1);
          ptx::mbar_arrive_expect_tx(&bar->v_full[s], KV_BYTES);
          ptx::tma_load_4d(sV + s * KV_BYTES, mv, &bar->v_full[s], 0, head, i * BN, fr);
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ==================================
    if (ptx::elect_one()) {
      constexpr int fmt = std::is_same<T, __nv_bfloat16>::value ? 1 : 0;
      constexpr uint32_t idesc_qk = ptx::make_idesc(BM, BN, fmt, 0);  // S = Q K^T : both operands K-major
      constexpr uint32_t idesc_pv = ptx::make_idesc(BM, D, fmt, 1);   // O += P V  : V is N(=d)-contiguous
      const uint32_t q_addr = ptx::smem_u32(sQ), k_addr = ptx::smem_u32(sK), v_addr = ptx::smem_u32(sV);
      auto issue_qk = [&](int t, int s) {
        const uint64_t qd = ptx::make_smem_desc_sw128(q_addr + t * Q_BYTES, 16, 1024);
        const uint64_t kd = ptx::make_smem_desc_sw128(k_addr + s * KV_BYTES, 16, 1024);
#pragma unroll
        for (int k = 0; k < D / 16; ++k)
          ptx::mma_ss(tmem + TMEM_S + t * BN, qd + 2 * k, kd + 2 * k, idesc_qk, k != 0);
      };
      ptx::mbar_wait(&bar->q_full, 0);
      ptx::mbar_wait(&bar->k_full[0], 0);
      ptx::tc_fence_after();
      for (int t = 0; t < QT; ++t) { issue_qk(t, 0); ptx::tc_commit(&bar->s_full[t]); }
      ptx::tc_commit(&bar->k_empty[0]);
      int j = 0;
      for (int slot = 0; slot < 3; ++slot) {
        for (int i = 0; i < nsteps[slot]; ++i, ++j) {
          const int s = j % ST, s1 = (j + 1) % ST;
          const bool more = j + 1 < total_steps;
          ptx::mbar_wait(&bar->v_full[s], (j / ST) & 1);
          if (more) ptx::mbar_wait(&bar->k_full[s1], ((j + 1) / ST) & 1);
          for (int t = 0; t < QT; ++t) {
            ptx::mbar_wait(&bar->p_full[t], j & 1);
            ptx::tc_fence_after();
            const uint32_t o_t = tmem + TMEM_O + (t * 3 + slot) * D;
            const uint32_t p_t = tmem + TMEM_S + t * BN;
#pragma unroll
            for (int k = 0; k < BN / 16; ++k) {
              // 16 keys per MMA: 8 packed columns of P, 16 rows (2048 B) of the V tile
              const uint64_t vd = ptx::make_smem_desc_sw128(v_addr + s * KV_BYTES + k * 2048, 16, 1024);
              ptx::mma_ts(o_t, p_t + k * 8, vd, idesc_pv, (i | k) != 0);
            }
            if (more) issue_qk(t, s1);
            ptx::tc_commit(&bar->s_full[t]);  // S_t(j+1) ready; on the last step: all of O_t ready
          }
          ptx::tc_commit(&bar->v_empty[s]);
          if (more) ptx::tc_commit(&bar->k_empty[s1]);
        }
      }
    }
  } else {
    // ================================ softmax warpgroups ==========================
    const int t = (warp - 2) >> 2;   // Q tile of this warpgroup
    const int quad = warp & 3;       // TMEM lane quadrant of this warp
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    const uint32_t s_addr = tmem + lane_base + TMEM_S + t * BN;
    const float sl2 = a.scale_log2;
    float m_slot[3] = {-INFINITY, -INFINITY, -INFINITY}, l_slot[3] = {0.f, 0.f, 0.f};
    int j = 0;
#pragma unroll
    for (int slot = 0; slot < 3; ++slot) {
      float m_ref = -INFINITY, l = 0.f;  // m_ref in raw-score units
      const uint32_t o_addr = tmem + lane_base + TMEM_O + (t * 3 + slot) * D;
      for (int i = 0; i < nsteps[slot]; ++i, ++j) {
        ptx::mbar_wait(&bar->s_full[t], j & 1);
        ptx::tc_fence_after();
        uint32_t sr[2][32];
        ptx::tmem_ld32(s_addr, sr[0]);
        ptx::tmem_ld32(s_addr + 32, sr[1]);
        ptx::tmem_wait_ld();
        const int valid = a.L - i * BN;  // keys of this tile that exist
        if (valid < BN) {
#pragma unroll
          for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int e = 0; e < 32; ++e)
              if (h * 32 + e >= valid) sr[h][e] = __float_as_uint(-INFINITY);
        }
        float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
        for (int e = 0; e < 32; e += 4) {
          mx0 = fmaxf(mx0, fmaxf(__uint_as_float(sr[0][e]), __uint_as_float(sr[1][e])));
          mx1 = fmaxf(mx1, fmaxf(__uint_as_float(sr[0][e + 1]), __uint_as_float(sr[1][e + 1])));
          mx2 = fmaxf(mx2, fmaxf(__uint_as_float(sr[0][e + 2]), __uint_as_float(sr[1][e + 2])));
          mx3 = fmaxf(mx3, fmaxf(__uint_as_float(sr[0][e + 3]), __uint_as_float(sr[1][e + 3])));
        }
        const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
        if (i == 0) {
          m_ref = mx;  // fresh accumulator: the first P.V of a slot overwrites O
        } else {
          const bool grow = (mx - m_ref) * sl2 > kRescaleThreshold;
          if (__any_sync(0xffffffffu, grow)) {
            // O_{t,slot} is quiescent here: s_full fired after every earlier MMA of this tile completed
            const float m_new = grow ? mx : m_ref;
            const float alpha = exp2f((m_ref - m_new) * sl2);
            l *= alpha;
            m_ref = m_new;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              uint32_t o[32];
              ptx::tmem_ld32(o_addr + h * 32, o);
              ptx::tmem_wait_ld();
#pragma unroll
              for (int e = 0; e < 32; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
              ptx::tmem_st32(o_addr + h * 32, o);
            }
          }
        }
        const float neg = -m_ref * sl2;
        uint32_t pk[32];
        float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int e = 0; e < 32; e += 2) {
            const float p0 = exp2f(fmaf(__uint_as_float(sr[h][e]), sl2, neg));
            const float p1 = exp2f(fmaf(__uint_as_float(sr[h][e + 1]), sl2, neg));
            sum0 += p0; sum1 += p1;
            pk[h * 16 + e / 2] = pack2<T>(p0, p1);
          }
        l += sum0 + sum1;
        ptx::tmem_st32(s_addr, pk);  // P_t over the low half of S_t
        ptx::tmem_wait_st();
        ptx::tc_fence_before();
        ptx::mbar_arrive(&bar->p_full[t]);
      }
      m_slot[slot] = m_ref * sl2;
      l_slot[slot] = l;
    }
    // ---- epilogue: merge the slots, alpha-lerp, write the head's 64 output channels of this row ----
    ptx::mbar_wait(&bar->s_full[t], j & 1);
    ptx::tc_fence_after();
    float cf[3];
    merge_coefficients(plan, m_slot[0], l_slot[0], m_slot[1], l_slot[1], m_slot[2], l_slot[2], cf[0], cf[1], cf[2]);
    const int row = row0 + t * BM + quad * 32 + lane;
    T* dst = (T*)a.out + ((long long)n * a.S + row) * (a.heads * D) + head * D;
    const bool used[3] = {plan.use0, plan.use1, plan.use2};
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float acc[32];
#pragma unroll
      for (int e = 0; e < 32; ++e) acc[e] = 0.f;
#pragma unroll
      for (int slot = 0; slot < 3; ++slot) {
        if (!used[slot]) continue;  // CTA-uniform
        uint32_t o[32];
        ptx::tmem_ld32(tmem + lane_base + TMEM_O + (t * 3 + slot) * D + h * 32, o);
        ptx::tmem_wait_ld();
#pragma unroll
        for (int e = 0; e < 32; ++e) acc[e] = fmaf(cf[slot], __uint_as_float(o[e]), acc[e]);
      }
      if (row < a.S) {
#pragma unroll
        for (int v = 0; v < 4; ++v)
          *reinterpret_cast<uint4*>(dst + h * 32 + v * 8) =
              make_uint4(pack2<T>(acc[v * 8], acc[v * 8 + 1]), pack2<T>(acc[v * 8 + 2], acc[v * 8 + 3]),
                         pack2<T>(acc[v * 8 + 4], acc[v * 8 + 5]), pack2<T>(acc[v * 8 + 6], acc[v * 8 + 7]));
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) { __syncwarp(); ptx::tmem_dealloc(tmem, 512); }
}

template <typename T>
int launch_t(const CUtensorMap* maps, const TcArgs& ta, cudaStream_t stream) {
  auto kern = attn_tc_kernel<T>;
  static bool configured = false;
  if (!configured) {
    PAID_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    configured = true;
  }
  dim3 grid((ta.S + QT * BM - 1) / (QT * BM), ta.heads, ta.N);
  kern<<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(maps[0], maps[1], maps[2], maps[3], maps[4], maps[5], maps[6], ta);
  PAID_LAUNCH_CHECK("attn_tc_kernel");
  return PAID_OK;
}

}  // namespace

bool attn_tc_supported(const CoreArgs& a) {
  return a.head_dim == D && a.heads <= 65535 && a.N <= 65535 &&
         !(((uintptr_t)a.q | (uintptr_t)a.k | (uintptr_t)a.v | (uintptr_t)a.out | (uintptr_t)a.k1 | (uintptr_t)a.v1 |
            (uintptr_t)a.k2 | (uintptr_t)a.v2) & 15);
}

int launch_attn_tc(const CoreArgs& a, cudaStream_t stream) {
  CUtensorMap maps[7];
  const long long C = (long long)a.heads * D;
  int st = make_tmap_heads(&maps[0], a.q, a.dtype, a.N, a.S, a.heads, D, (long long)a.S * C, BM);
  if (st != PAID_OK) return st;
  if ((st = make_tmap_heads(&maps[1], a.k, a.dtype, a.N, a.L, a.heads, D, (long long)a.L * C, BN)) != PAID_OK) return st;
  if ((st = make_tmap_heads(&maps[2], a.v, a.dtype, a.N, a.L, a.heads, D, (long long)a.L * C, BN)) != PAID_OK) return st;
  TcArgs ta{};
  ta.per_frame[0] = 1;
  const void* kk[2] = {a.k1, a.k2};
  const void* vv[2] = {a.v1, a.v2};
  const long long strides[2] = {a.stride1, a.stride2};
  for (int s = 0; s < 2; ++s) {
    if (kk[s]) {
      const long long frames = strides[s] ? a.N : 1;
      if ((st = make_tmap_heads(&maps[3 + 2 * s], kk[s], a.dtype, frames, a.L, a.heads, D, strides[s], BN)) != PAID_OK) return st;
      if ((st = make_tmap_heads(&maps[4 + 2 * s], vv[s], a.dtype, frames, a.L, a.heads, D, strides[s], BN)) != PAID_OK) return st;
      ta.per_frame[1 + s] = strides[s] ? 1 : 0;
    } else {  // unused slot: any valid descriptor
      maps[3 + 2 * s] = maps[1];
      maps[4 + 2 * s] = maps[2];
      ta.per_frame[1 + s] = 1;
    }
  }
  ta.mode = a.mode; ta.fused = a.fused; ta.N = a.N; ta.S = a.S; ta.L = a.L; ta.heads = a.heads;
  ta.begin_frame = a.begin_frame; ta.end_frame = a.end_frame;
  ta.scale_log2 = a.scale * kLog2e;
  ta.coef = a.coef; ta.out = a.out;
  return a.dtype == PAID_F16 ? launch_t<__half>(maps, ta, stream) : launch_t<__nv_bfloat16>(maps, ta, stream);
}

}  // namespace paid
