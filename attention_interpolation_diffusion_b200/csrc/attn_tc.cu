// Interpolated attention core on the 5th-generation tensor cores: head_dim 64 (SDXL), and any multiple of 8 up to 192
// as DT = ceil(head_dim / 64) chunks of 64 columns, the last one zero-padded by the TMA unit (SD1.5: 40, 80, 160).
//
// Replaces, per attention layer, the reference's endpoint replication, head split, [self ; endpoint]
// concatenations, the two materialised softmax(QK^T) matrices, the two P.V products and the alpha-lerp
// (interpolation.py:627-664 outer, 760-790 inner; deactivated 581-584) with ONE kernel launch.
//
// Formulation.  The output of frame n is  wA * softmax-attn(q, keys_A) + wB * softmax-attn(q, keys_B)  with
//   stream A keys = [own K/V (if fused) ; endpoint slot 1],   stream B keys = [own K/V (if fused) ; slot 2]
// (paid_common.cuh FramePlan).  Each stream is one flash-style online softmax with its own fp32 accumulator
// in TMEM.  The shared "own K/V" segment is scored and exponentiated ONCE (both streams have identical
// running statistics while they have seen the same keys) and its P.V product is issued into both
// accumulators; endpoint K/V are read once per CTA and never replicated or concatenated.
//
// Work decomposition: one CTA per (frame, head, QT x 128 query rows); QT Q tiles share every K/V tile.
//   warp 0        TMA producer: Q tile(s) once, then K / V tiles (64 keys) into a 5- or 8-stage shared-memory ring
//   warp 1        tcgen05.mma issuer: S_t = Q_t K^T (SS), acc_{t,stream} += P_t V (A operand P from TMEM)
//   warps 2-3     idle (they only give their registers away)
//   warps 4-7     softmax warpgroup of Q tile 0 (one thread per query row)
//   warps 8-11    softmax warpgroup of Q tile 1 (QT = 2 only)
// The score tile of each Q tile is DOUBLE-BUFFERED in TMEM: S_t(j+2) = Q_t K(j+2)^T is issued right after
// P_t(j) V(j), so the scores of the next step are ready while the softmax warps still work on the current
// one and the exp-bound softmax never waits for the tensor core.
// TMEM (256 QT columns): S_{t,b} at 64 (2 t + b) (P_{t,b} aliases its low 32 columns as packed 16-bit),
// acc_{t,stream} at 128 QT + 128 t + 64 stream.
#include <cstdlib>
#include <type_traits>

#include "paid_common.cuh"
#include "sm100_ptx.cuh"

namespace paid {
namespace {

constexpr int D = 64;            // head_dim of the tiles.  A smaller head_dim (multiple of 8) runs zero-padded: the 4-D
                                 // tensor maps have extent head_dim in their innermost dimension and a 64-wide box, so
                                 // the TMA unit fills columns head_dim..63 of every Q/K/V tile with zeros (scores and
                                 // the first head_dim accumulator columns are exact); the epilogue stores head_dim columns.
constexpr int BM = 128;          // rows per Q tile
// Q tiles per CTA is a template parameter.  QT = 1 (default): one 128-row Q tile, half the TMEM / shared memory /
// threads, TWO CTAs per SM: the two resident CTAs run out of phase, so one computes while the other starts, waits
// for the tensor core or drains (measured 6-11 % faster than QT = 2 on every SDXL shape).  QT = 2 (PAID_ATTN_QT=2):
// one CTA per SM whose two Q tiles share every K/V tile (half the K/V traffic from L2 to shared memory).
constexpr int BN = 64;           // keys per step
constexpr int ST = 8;            // K/V ring stages
constexpr int Q_BYTES = BM * D * 2;    // 16 KB: one 64-column chunk of a Q tile
constexpr int KV_BYTES = BN * D * 2;   //  8 KB: one 64-column chunk of a K or V tile
// DT = 64-column chunks of the head dimension (1: head_dim <= 64, 2: <= 128, 3: <= 192).  A chunk is one 128-byte
// swizzled shared-memory tile per operand: Q K^T accumulates over the chunks (K dimension), P V is issued once per
// chunk (N = 64 each) into adjacent accumulator columns, so every MMA and descriptor is the head_dim-64 one.
// DT > 1 needs 128 + 2 * 64 DT > 256 TMEM columns: one CTA per SM, QT = 1 only.
template <int QT, int DT> struct Shape {
  static_assert(DT == 1 || QT == 1, "wide heads run with one Q tile per CTA");
  static constexpr int kStages = DT == 3 ? 3 : (QT == 2 ? ST : 5);   // K/V ring stages (227 KB per SM)
  static constexpr int kSmemBytes = 1024 + QT * DT * Q_BYTES + kStages * 2 * DT * KV_BYTES + 512;
  static constexpr int kThreads = 128 * (1 + QT);   // warpgroup 0: TMA + MMA (+2 idle warps); the others: softmax
  static constexpr int kCtasPerSm = (QT == 2 || DT > 1) ? 1 : 2;
  // setmaxnreg split of the register file among the CTA's warpgroups (per-CTA budget 64K / kCtasPerSm)
  // setmaxnreg.inc draws from the CTA's OWN launch allocation (kThreads x launch registers), so the split must satisfy
  // 128 kRegsControl + (kThreads - 128) kRegsSoftmax <= kThreads x launch registers, or the .inc never returns.
  static constexpr int kRegsControl = 56;
  static constexpr int kRegsSoftmax = QT == 2 ? 224 : 200;
  static constexpr uint32_t kTmemCols = DT == 1 ? 256 * QT : 512;
  static constexpr uint32_t kTmemAcc = 128 * QT;     // S buffers first (2 x 64 columns per tile), accumulators after
  static constexpr uint32_t kAccTile = 128 * DT;     // accumulator columns per Q tile: 2 streams x 64 DT
};
constexpr uint32_t TMEM_S = 0;
constexpr float kRescaleThreshold = 8.f;  // log2 units: rescale an accumulator only when its max grew by > 2^8

struct TcArgs {
  int mode, fused, N, S, L, heads, head_dim, begin_frame, end_frame;
  float scale_log2;
  const float* coef;
  void* out;
  int accumulate;
  float out_scale;
  const float* out_frame_scale;
  int per_frame[3];  // slot K/V map has one matrix per frame (1) or a single shared matrix (0)
  int wide;          // output rows of a head start on 32-byte boundaries: 32-byte stores
};

struct Barriers {
  uint64_t q_full;
  uint64_t k_full[ST], k_empty[ST], v_full[ST], v_empty[ST];
  uint64_t s_full[2][2], p_full[2][2], pv_done[2], acc_final[2];
  uint32_t tmem_slot;
};

// key segments of one frame: K/V slot and the streams it feeds (bit 0: A, bit 1: B)
struct Segments {
  int count;
  int slot[3];
  int feeds[3];
  bool a_active, b_active;
};

__device__ __forceinline__ Segments make_segments(const FramePlan& p) {
  Segments g;
  g.count = 0;
  g.a_active = p.wA != 0.f;
  g.b_active = p.wB != 0.f;
  if (p.use0) { g.slot[g.count] = 0; g.feeds[g.count] = (g.a_active ? 1 : 0) | (g.b_active ? 2 : 0); ++g.count; }
  if (p.use1) { g.slot[g.count] = 1; g.feeds[g.count] = 1; ++g.count; }
  if (p.use2) { g.slot[g.count] = 2; g.feeds[g.count] = 2; ++g.count; }
  return g;
}

__device__ __forceinline__ int frame_of_block(int z, int N) {
  // heavy (interior) frames first, the two cheap endpoint frames last
  if (N < 3) return z;
  return z < N - 2 ? z + 1 : (z == N - 2 ? 0 : N - 1);
}

template <typename T, int QT, int DT>
__global__ void __launch_bounds__(Shape<QT, DT>::kThreads, Shape<QT, DT>::kCtasPerSm)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK0,
               const __grid_constant__ CUtensorMap tmV0, const __grid_constant__ CUtensorMap tmK1,
               const __grid_constant__ CUtensorMap tmV1, const __grid_constant__ CUtensorMap tmK2,
               const __grid_constant__ CUtensorMap tmV2, const TcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  using SH = Shape<QT, DT>;
  constexpr int ST = SH::kStages;                  // (shadows the namespace constant: ring depth of this variant)
  constexpr uint32_t TMEM_ACC = SH::kTmemAcc;
  constexpr int QTILE = DT * Q_BYTES, KVTILE = DT * KV_BYTES;   // bytes of one Q / K / V tile (DT chunks)
  uint8_t* sQ = smem;                              // [QT][DT][128][64]
  uint8_t* sK = sQ + QT * QTILE;                   // [ST][DT][64][64]
  uint8_t* sV = sK + ST * KVTILE;                  // [ST][DT][64][64]
  Barriers* bar = reinterpret_cast<Barriers*>(sV + ST * KVTILE);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = frame_of_block(blockIdx.z, a.N), head = blockIdx.y, row0 = blockIdx.x * (QT * BM);

  if (warp == 0 && ptx::elect_one()) {
    ptx::prefetch_tmap(&tmQ);
    ptx::prefetch_tmap(&tmK0); ptx::prefetch_tmap(&tmV0);
    ptx::mbar_init(&bar->q_full, 1);
    for (int s = 0; s < ST; ++s) {
      ptx::mbar_init(&bar->k_full[s], 1); ptx::mbar_init(&bar->k_empty[s], 1);
      ptx::mbar_init(&bar->v_full[s], 1); ptx::mbar_init(&bar->v_empty[s], 1);
    }
    for (int t = 0; t < QT; ++t) {
      for (int b = 0; b < 2; ++b) {
        ptx::mbar_init(&bar->s_full[t][b], 1);
        ptx::mbar_init(&bar->p_full[t][b], 4);  // one arrive per softmax warp
      }
      ptx::mbar_init(&bar->pv_done[t], 1);
      ptx::mbar_init(&bar->acc_final[t], 1);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) { ptx::tmem_alloc(&bar->tmem_slot, SH::kTmemCols); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = bar->tmem_slot;
  ptx::pdl_launch_dependents();  // the next kernel may begin its prologue
  ptx::pdl_wait();               // everything above overlapped the previous kernel's tail; its results are visible now

  const float c = a.mode == PAID_PLAIN ? 0.f : a.coef[n];
  const FramePlan plan = make_frame_plan(a.mode, a.fused, n, a.begin_frame, a.end_frame, c);
  const Segments seg = make_segments(plan);
  const int tiles = (a.L + BN - 1) / BN;
  const int total_steps = seg.count * tiles;

  if (warp < 4) {
  ptx::setmaxnreg_dec<SH::kRegsControl>();
  if (warp == 0) {
    // ================================ TMA producer ================================
    if (ptx::elect_one()) {
      ptx::mbar_arrive_expect_tx(&bar->q_full, QT * QTILE);
      for (int t = 0; t < QT; ++t)
        for (int c = 0; c < DT; ++c)
          ptx::tma_load_4d(sQ + t * QTILE + c * Q_BYTES, &tmQ, &bar->q_full, c * D, head, row0 + t * BM, n);
      int j = 0;
      for (int g = 0; g < seg.count; ++g) {
        const int slot = seg.slot[g];
        const CUtensorMap* mk = slot == 0 ? &tmK0 : (slot == 1 ? &tmK1 : &tmK2);
        const CUtensorMap* mv = slot == 0 ? &tmV0 : (slot == 1 ? &tmV1 : &tmV2);
        const int fr = a.per_frame[slot] ? n : 0;
        for (int i = 0; i < tiles; ++i, ++j) {
          const int s = j % ST;
          const uint32_t ph = (j / ST) & 1;
          ptx::mbar_wait(&bar->k_empty[s], ph ^ 1);
          ptx::mbar_arrive_expect_tx(&bar->k_full[s], KVTILE);
          for (int c = 0; c < DT; ++c)
            ptx::tma_load_4d(sK + s * KVTILE + c * KV_BYTES, mk, &bar->k_full[s], c * D, head, i * BN, fr);
          ptx::mbar_wait(&bar->v_empty[s], ph ^ 1);
          ptx::mbar_arrive_expect_tx(&bar->v_full[s], KVTILE);
          for (int c = 0; c < DT; ++c)
            ptx::tma_load_4d(sV + s * KVTILE + c * KV_BYTES, mv, &bar->v_full[s], c * D, head, i * BN, fr);
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ==================================
    if (ptx::elect_one()) {
      constexpr int fmt = std::is_same<T, __nv_bfloat16>::value ? 1 : 0;
      constexpr uint32_t idesc_qk = ptx::make_idesc(BM, BN, fmt, 0);  // S = Q K^T : both operands K-major
      constexpr uint32_t idesc_pv = ptx::make_idesc(BM, D, fmt, 1);   // acc += P V : V is N(=d)-contiguous
      const uint32_t q_addr = ptx::smem_u32(sQ), k_addr = ptx::smem_u32(sK), v_addr = ptx::smem_u32(sV);
      auto issue_qk = [&](int t, int b, int s) {
#pragma unroll
        for (int c = 0; c < DT; ++c) {   // the head dimension is the K dimension of this product: accumulate over chunks
          const uint64_t qd = ptx::make_smem_desc_sw128(q_addr + t * QTILE + c * Q_BYTES, 16, 1024);
          const uint64_t kd = ptx::make_smem_desc_sw128(k_addr + s * KVTILE + c * KV_BYTES, 16, 1024);
#pragma unroll
          for (int k = 0; k < D / 16; ++k)
            ptx::mma_ss(tmem + TMEM_S + (t * 2 + b) * BN, qd + 2 * k, kd + 2 * k, idesc_qk, (c | k) != 0);
        }
      };
      // prologue: scores of steps 0 and 1
      ptx::mbar_wait(&bar->q_full, 0);
      for (int jj = 0; jj < 2 && jj < total_steps; ++jj) {
        ptx::mbar_wait(&bar->k_full[jj], 0);
        ptx::tc_fence_after();
        for (int t = 0; t < QT; ++t) { issue_qk(t, jj, jj); ptx::tc_commit(&bar->s_full[t][jj]); }
        ptx::tc_commit(&bar->k_empty[jj]);
      }
      bool started[2] = {false, false};  // has stream A / B received a P.V product yet
      int j = 0;
      for (int g = 0; g < seg.count; ++g) {
        const int feeds = seg.feeds[g];
        for (int i = 0; i < tiles; ++i, ++j) {
          const int s = j % ST, s2 = (j + 2) % ST, b = j & 1;
          const bool more = j + 2 < total_steps;
          ptx::mbar_wait(&bar->v_full[s], (j / ST) & 1);
          if (more) ptx::mbar_wait(&bar->k_full[s2], ((j + 2) / ST) & 1);
          for (int t = 0; t < QT; ++t) {
            ptx::mbar_wait(&bar->p_full[t][b], (j >> 1) & 1);
            ptx::tc_fence_after();
            const uint32_t p_t = tmem + TMEM_S + (t * 2 + b) * BN;
#pragma unroll
            for (int st = 0; st < 2; ++st) {
              if (!(feeds & (1 << st))) continue;
#pragma unroll
              for (int c = 0; c < DT; ++c) {   // one N = 64 product per chunk of the head dimension
                const uint32_t acc = tmem + TMEM_ACC + t * SH::kAccTile + (st * DT + c) * D;  // 2 streams x DT x 64 columns
#pragma unroll
                for (int k = 0; k < BN / 16; ++k) {
                  // 16 keys per MMA: 8 packed columns of P, 16 rows (2048 B) of the V chunk
                  const uint64_t vd = ptx::make_smem_desc_sw128(v_addr + s * KVTILE + c * KV_BYTES + k * 2048, 16, 1024);
                  ptx::mma_ts(acc, p_t + k * 8, vd, idesc_pv, started[st] || k != 0);
                }
              }
            }
            ptx::tc_commit(&bar->pv_done[t]);     // accumulators of tile t include step j
            if (j + 1 == total_steps) ptx::tc_commit(&bar->acc_final[t]);  // single-use: everything has landed
            if (more) {                            // S_{t,b} is free again (in-order after the P.V above)
              issue_qk(t, b, s2);
              ptx::tc_commit(&bar->s_full[t][b]);  // scores of step j + 2
            }
          }
          if (feeds & 1) started[0] = true;
          if (feeds & 2) started[1] = true;
          ptx::tc_commit(&bar->v_empty[s]);
          if (more) ptx::tc_commit(&bar->k_empty[s2]);
        }
      }
    }
  }
  } else {
    ptx::setmaxnreg_inc<SH::kRegsSoftmax>();
    // ================================ softmax warpgroups ==========================
    const int t = (warp - 4) >> 2;   // Q tile of this warpgroup
    const int quad = warp & 3;       // TMEM lane quadrant of this warp
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    const uint32_t s_base = tmem + lane_base + TMEM_S + t * 2 * BN;
    const uint32_t acc_addr = tmem + lane_base + TMEM_ACC + t * SH::kAccTile;
    const float sl2 = a.scale_log2;
    float m_st[2] = {-INFINITY, -INFINITY}, l_st[2] = {0.f, 0.f};  // per stream, m in raw-score units
    bool started[2] = {false, false};
    int j = 0;
    for (int g = 0; g < seg.count; ++g) {
      const int feeds = seg.feeds[g];
      const int primary = (feeds & 1) ? 0 : 1;   // streams fed together have identical statistics
      float m_ref = m_st[primary], l = l_st[primary];
      const bool fresh = !started[primary];
      for (int i = 0; i < tiles; ++i, ++j) {
        const int b = j & 1;
        const uint32_t s_addr = s_base + b * BN;
        ptx::mbar_wait(&bar->s_full[t][b], (j >> 1) & 1);
        ptx::tc_fence_after();
        uint32_t sr[2][32];
#pragma unroll
        for (int h = 0; h < 2; ++h) ptx::tmem_ld32(s_addr + h * 32, sr[h]);
        ptx::tmem_wait_ld();
        const int valid = a.L - i * BN;  // keys of this tile that exist
        if (valid < BN) {                // ragged last tile only (kept a real branch by the asm statement)
          asm volatile("" ::: "memory");
#pragma unroll
          for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int e = 0; e < 32; ++e)
              if (h * 32 + e >= valid) sr[h][e] = __float_as_uint(-INFINITY);
        }
        float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          mx0 = fmaxf(mx0, __uint_as_float(sr[0][e]));
          mx1 = fmaxf(mx1, __uint_as_float(sr[0][e + 1]));
          mx2 = fmaxf(mx2, __uint_as_float(sr[1][e]));
          mx3 = fmaxf(mx3, __uint_as_float(sr[1][e + 1]));
        }
        const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
        if (i == 0 && fresh) {
          m_ref = mx;  // fresh accumulators: the first P.V of a stream overwrites them
        } else {
          const bool grow = (mx - m_ref) * sl2 > kRescaleThreshold;
          if (__any_sync(0xffffffffu, grow)) {
            // rare: wait until the P.V of step j-1 has landed, then the accumulators are quiescent (the P.V of
            // step j cannot be issued before this thread publishes P(j))
            ptx::mbar_wait(&bar->pv_done[t], (j - 1) & 1);
            ptx::tc_fence_after();
            const float m_new = grow ? mx : m_ref;
            const float alpha = ptx::ex2((m_ref - m_new) * sl2);
            l *= alpha;
            m_ref = m_new;
#pragma unroll
            for (int st = 0; st < 2; ++st) {
              if (!(feeds & (1 << st))) continue;
#pragma unroll
              for (int h = 0; h < 2 * DT; ++h) {
                uint32_t o[32];
                ptx::tmem_ld32(acc_addr + st * DT * D + h * 32, o);
                ptx::tmem_wait_ld();
#pragma unroll
                for (int e = 0; e < 32; e += 2) {
                  const float2 r = ptx::mul2(make_float2(__uint_as_float(o[e]), __uint_as_float(o[e + 1])),
                                             make_float2(alpha, alpha));
                  o[e] = __float_as_uint(r.x); o[e + 1] = __float_as_uint(r.y);
                }
                ptx::tmem_st32(acc_addr + st * DT * D + h * 32, o);
              }
            }
          }
        }
        const float neg = -m_ref * sl2;
        const float2 sl2v = make_float2(sl2, sl2), negv = make_float2(neg, neg);
        float2 sumA = make_float2(0.f, 0.f), sumB = make_float2(0.f, 0.f);
        // Three phases over the 64 scores, each a run of independent instructions: packed FMAs (x * scale*log2e -
        // max), then the MUFU.EX2 burst, then packed adds / converts.  While this warp sits in its SFU burst the
        // other softmax warp of the scheduler runs its FMA/ALU phases, so the SFU stays busy.
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int e = 0; e < 32; e += 2) {
            const float2 x = ptx::fma2(make_float2(__uint_as_float(sr[h][e]), __uint_as_float(sr[h][e + 1])), sl2v, negv);
            sr[h][e] = __float_as_uint(x.x); sr[h][e + 1] = __float_as_uint(x.y);
          }
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int e = 0; e < 32; ++e) sr[h][e] = __float_as_uint(ptx::ex2v(__uint_as_float(sr[h][e])));
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t pk[16];
#pragma unroll
          for (int e = 0; e < 32; e += 4) {
            const float2 x0 = make_float2(__uint_as_float(sr[h][e]), __uint_as_float(sr[h][e + 1]));
            const float2 x1 = make_float2(__uint_as_float(sr[h][e + 2]), __uint_as_float(sr[h][e + 3]));
            sumA = ptx::add2(sumA, x0);
            sumB = ptx::add2(sumB, x1);
            pk[e / 2] = pack2<T>(x0.x, x0.y);
            pk[e / 2 + 1] = pack2<T>(x1.x, x1.y);
          }
          ptx::tmem_st16(s_addr + h * 16, pk);  // P over the low half of this S buffer (all of S is in registers)
        }
        l += (sumA.x + sumA.y) + (sumB.x + sumB.y);
        ptx::tmem_wait_st();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&bar->p_full[t][b]);
      }
#pragma unroll
      for (int st = 0; st < 2; ++st)
        if (feeds & (1 << st)) { m_st[st] = m_ref; l_st[st] = l; started[st] = true; }
    }
    // ---- epilogue: out = wA * accA / lA + wB * accB / lB for this row's 64 channels of the head ----
    // All P.V products of this tile must have landed: a dedicated single-phase barrier (pv_done completes once
    // per step, and a parity wait cannot be used on a barrier whose phases this thread has skipped).
    ptx::mbar_wait(&bar->acc_final[t], 0);
    // ... and pv_done has completed its last phase by now (committed just before acc_final by the same thread): observing
    // it costs one successful poll and leaves no barrier phase unwaited at exit (compute-sanitizer synccheck "Missing wait")
    ptx::mbar_wait(&bar->pv_done[t], (total_steps - 1) & 1);
    ptx::tc_fence_after();
    const float os = a.out_scale * (a.out_frame_scale ? a.out_frame_scale[n] : 1.f);
    const float cf[2] = {seg.a_active ? os * plan.wA / l_st[0] : 0.f, seg.b_active ? os * plan.wB / l_st[1] : 0.f};
    const bool active[2] = {seg.a_active, seg.b_active};
    const int row = row0 + t * BM + quad * 32 + lane;
    T* dst = (T*)a.out + ((long long)n * a.S + row) * (a.heads * a.head_dim) + head * a.head_dim;
#pragma unroll
    for (int h = 0; h < 2 * DT; ++h) {
      if (h * 32 >= a.head_dim) break;   // padded columns
      float acc[32];
#pragma unroll
      for (int e = 0; e < 32; ++e) acc[e] = 0.f;
#pragma unroll
      for (int st = 0; st < 2; ++st) {
        if (!active[st]) continue;  // CTA-uniform
        uint32_t o[32];
        ptx::tmem_ld32(acc_addr + st * DT * D + h * 32, o);
        ptx::tmem_wait_ld();
#pragma unroll
        for (int e = 0; e < 32; ++e) acc[e] = fmaf(cf[st], __uint_as_float(o[e]), acc[e]);
      }
      if (row < a.S) {
        if (a.accumulate) {  // out += ...: the IP-Adapter second attention (CTA-uniform branch)
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            if (h * 32 + v * 8 >= a.head_dim) break;
            const uint4 old = *reinterpret_cast<const uint4*>(dst + h * 32 + v * 8);
            const T* o8 = reinterpret_cast<const T*>(&old);
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[v * 8 + e] += to_f32(o8[e]);
          }
        }
#pragma unroll
        for (int v = 0; v < 2; ++v) {   // 16 channels per step; padded head_dim: columns head_dim..63 are zero and are not stored
          if (h * 32 + v * 16 >= a.head_dim) break;
          uint32_t w[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) w[e] = pack2<T>(acc[v * 16 + 2 * e], acc[v * 16 + 2 * e + 1]);
          const bool both = h * 32 + v * 16 + 8 < a.head_dim;   // head_dim % 16 == 8: the last step has 8 channels
          if (a.wide && both) {
            ptx::st_global_256(dst + h * 32 + v * 16, w);
          } else {
            *reinterpret_cast<uint4*>(dst + h * 32 + v * 16) = make_uint4(w[0], w[1], w[2], w[3]);
            if (both) *reinterpret_cast<uint4*>(dst + h * 32 + v * 16 + 8) = make_uint4(w[4], w[5], w[6], w[7]);
          }
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) { __syncwarp(); ptx::tmem_dealloc(tmem, SH::kTmemCols); }
}

template <typename T, int QT, int DT>
int launch_t(const CUtensorMap* maps, const TcArgs& ta, cudaStream_t stream) {
  auto kern = attn_tc_kernel<T, QT, DT>;
  constexpr int SMEM_BYTES = Shape<QT, DT>::kSmemBytes, NUM_THREADS = Shape<QT, DT>::kThreads;
  static_assert(SMEM_BYTES <= 232448, "shared memory per CTA");
  PAID_CUDA_CHECK(ensure_kernel_configured((const void*)kern, SMEM_BYTES, nullptr));
  dim3 grid((ta.S + QT * BM - 1) / (QT * BM), ta.heads, ta.N);
  PAID_CUDA_CHECK(launch_pdl(kern, grid, dim3(NUM_THREADS), SMEM_BYTES, stream, maps[0], maps[1], maps[2], maps[3], maps[4],
                             maps[5], maps[6], ta));
  PAID_LAUNCH_CHECK("attn_tc_kernel");
  return PAID_OK;
}

}  // namespace

bool attn_tc_supported(const CoreArgs& a) {
  // debugging knobs, read once: PAID_ATTN_NO_PAD=1 serves head_dim != 64 with the generic kernel,
  // PAID_ATTN_MAX_TC_HEAD_DIM sends larger head_dim there
  static const char* nopad = getenv("PAID_ATTN_NO_PAD");
  static const char* cap = getenv("PAID_ATTN_MAX_TC_HEAD_DIM");
  static const int max_hd = cap ? atoi(cap) : 3 * D;
  const bool padded_ok = a.head_dim != D && a.head_dim >= 16 && a.head_dim <= 3 * D && a.head_dim <= max_hd &&
                         a.head_dim % 8 == 0 && !(nopad && nopad[0] == '1');
  return (a.head_dim == D || padded_ok) && a.heads <= 65535 && a.N <= 65535 &&
         !(((uintptr_t)a.q | (uintptr_t)a.k | (uintptr_t)a.v | (uintptr_t)a.out | (uintptr_t)a.k1 | (uintptr_t)a.v1 |
            (uintptr_t)a.k2 | (uintptr_t)a.v2) & 15);
}

int launch_attn_tc(const CoreArgs& a, cudaStream_t stream) {
  CUtensorMap maps[7];
  const int hd = a.head_dim;  // the box of every map is D = 64 wide; the TMA unit zero-fills columns >= hd
  const long long C = (long long)a.heads * hd;
  int st = make_tmap_heads(&maps[0], a.q, a.dtype, a.N, a.S, a.heads, hd, (long long)a.S * C, BM);
  // a driver that rejects a box wider than the tensor: nothing was launched, the caller falls back (core_dispatch)
  if (st != PAID_OK) return hd % D ? PAID_EUNSUPPORTED : st;
  const long long kv_frames = a.stride0 ? a.N : 1;
  if ((st = make_tmap_heads(&maps[1], a.k, a.dtype, kv_frames, a.L, a.heads, hd, a.stride0, BN)) != PAID_OK) return st;
  if ((st = make_tmap_heads(&maps[2], a.v, a.dtype, kv_frames, a.L, a.heads, hd, a.stride0, BN)) != PAID_OK) return st;
  TcArgs ta{};
  ta.per_frame[0] = a.stride0 ? 1 : 0;
  const void* kk[2] = {a.k1, a.k2};
  const void* vv[2] = {a.v1, a.v2};
  const long long strides[2] = {a.stride1, a.stride2};
  for (int s = 0; s < 2; ++s) {
    if (kk[s]) {
      const long long frames = strides[s] ? a.N : 1;
      if ((st = make_tmap_heads(&maps[3 + 2 * s], kk[s], a.dtype, frames, a.L, a.heads, hd, strides[s], BN)) != PAID_OK) return st;
      if ((st = make_tmap_heads(&maps[4 + 2 * s], vv[s], a.dtype, frames, a.L, a.heads, hd, strides[s], BN)) != PAID_OK) return st;
      ta.per_frame[1 + s] = strides[s] ? 1 : 0;
    } else {  // unused slot: any valid descriptor
      maps[3 + 2 * s] = maps[1];
      maps[4 + 2 * s] = maps[2];
      ta.per_frame[1 + s] = 1;
    }
  }
  ta.mode = a.mode; ta.fused = a.fused; ta.N = a.N; ta.S = a.S; ta.L = a.L; ta.heads = a.heads; ta.head_dim = hd;
  ta.begin_frame = a.begin_frame; ta.end_frame = a.end_frame;
  ta.scale_log2 = a.scale * kLog2e;
  ta.coef = a.coef; ta.out = a.out;
  ta.accumulate = a.accumulate; ta.out_scale = a.out_scale; ta.out_frame_scale = a.out_frame_scale;
  ta.wide = a.head_dim % 16 == 0 && !((uintptr_t)a.out & 31);
  const char* force = getenv("PAID_ATTN_QT");   // tests flip this between calls: not cached
  const bool single_tile = !(force && force[0] == '2');
  if (hd > 2 * D)
    return a.dtype == PAID_F16 ? launch_t<__half, 1, 3>(maps, ta, stream) : launch_t<__nv_bfloat16, 1, 3>(maps, ta, stream);
  if (hd > D)
    return a.dtype == PAID_F16 ? launch_t<__half, 1, 2>(maps, ta, stream) : launch_t<__nv_bfloat16, 1, 2>(maps, ta, stream);
  if (single_tile)
    return a.dtype == PAID_F16 ? launch_t<__half, 1, 1>(maps, ta, stream) : launch_t<__nv_bfloat16, 1, 1>(maps, ta, stream);
  return a.dtype == PAID_F16 ? launch_t<__half, 2, 1>(maps, ta, stream) : launch_t<__nv_bfloat16, 2, 1>(maps, ta, stream);
}

}  // namespace paid
