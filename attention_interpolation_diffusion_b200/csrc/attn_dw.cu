// Single-stream attention core (PLAIN = deactivated processor / stock attention, INNER = lerped endpoint K/V) on the
// 5th-generation tensor cores, head_dim <= 64: a PERSISTENT kernel with TWO independent softmax warpgroups per CTA, each
// owning one 128-row Q block of a 256-row item; the K / V tiles are staged once and feed both.
//
// Replaces the reference's stock attention of the deactivated processors (interpolation.py:581-584, 715-718: 75 of the
// 100 UNet forwards of a sequence) and the inner-interpolated attention (interpolation.py:760-790).
//
// Why this shape.  The one-warpgroup kernel (attn_tc.cu) left the SFU pipe 35 % idle: a softmax warp spends ~40 % of a key
// tile in its MUFU.EX2 burst and the rest in latencies around it (barrier polls, TMEM load / store round trips), and two
// such warps per scheduler cannot cover each other.  Two warpgroups per CTA (two CTAs per SM) put four softmax warps on
// every scheduler.  The first dual-warpgroup version of this file split the KEY tiles of one Q block between the
// warpgroups by parity and merged the two partial attentions in the epilogue: the exchange cost ~3500 cycles per item
// (15 % of a 16-tile item, and most of an L = 77 item: profiles/r2_dw_cycle_trace.log).  Here the warpgroups never
// meet: warpgroup b runs a complete flash-attention stream (all key tiles) for Q block b of the item, with its own
// MMA-issuing warp, score buffer, accumulator, running (max, sum) and epilogue.
//
//   warp 0        TMA producer: the two Q blocks of an item, K / V tiles (64 keys) through a 4-stage ring
//   warps 1, 2    tcgen05.mma issuers, one per warpgroup: S_b = Q_b K_j^T (SS), acc_b += P_b V_j (A operand P from TMEM)
//   warp 3        idle (gives its registers away)
//   warps 4-7     softmax warpgroup 0 (Q block 0 of the item), one thread per query row
//   warps 8-11    softmax warpgroup 1 (Q block 1)
// TMEM (256 columns): S_0 | S_1 | acc_0 | acc_1, 64 columns each; P_b overwrites the low 32 columns of S_b (packed 16-bit).
//
// Persistent: the grid is (at most) two CTAs per SM; CTA c walks the items c, c + grid, ... of the list
// (frame, head, 256-row Q block pair), Q block pair fastest.  The producer and the issuers run ahead across item
// boundaries (the next Q block lands while the softmax warps work on the last tile and the epilogue of the current one),
// so the per-CTA prologue (barrier init, TMEM allocation, descriptor fetch, first-load latency) is paid once per CTA.
//
// Speculative reference maximum: a tile is exponentiated against the running reference m_ref BEFORE its own maximum is
// known (the maximum is reduced alongside); only if some row's maximum exceeds m_ref by more than 2^8 (never, after the
// first tile, for real attention logits) is the accumulator rescaled and the tile redone.
//
// Exponentials on two pipes.  At head_dim 64 a 128 x 64 score tile costs the tensor pipe 256 cycles (Q K^T + P V) and the
// SFU 512 (16 MUFU.EX2 per clock per SM): the kernel is SFU-bound at half of the tensor peak.  kPolyPairs of every 16
// element pairs are therefore exponentiated on the FMA pipe instead (Cody-Waite: round to nearest integer with the
// 1.5 * 2^23 trick, cubic minimax polynomial of 2^f on [-0.5, 0.5], relative error 7.5e-5 -- below the 16-bit rounding of
// P, 4.9e-4 / 3.9e-3 -- exponent added in the integer domain), as packed f32x2 instructions.
#include <cstdlib>
#include <type_traits>

#include "paid_common.cuh"
#include "sm100_ptx.cuh"

// PAID_DW_TRACE: cycle accounting of CTA 0 (debug builds only, tools/build_variant.sh): prints where the issuer threads and
// one softmax thread per warpgroup spend their time
#ifdef PAID_DW_TRACE
#define TR_DECL(...) long long __VA_ARGS__
#define TR_T(v) const long long v = clock64()
#define TR_ADD(acc, a, b) acc += (b) - (a)
#else
#define TR_DECL(...)
#define TR_T(v)
#define TR_ADD(acc, a, b)
#endif

#ifndef PAID_DW_POLY_PAIRS
#define PAID_DW_POLY_PAIRS 4   // of every 16 element pairs: 25 % of the exponentials leave the SFU
#endif

namespace paid {
namespace {

constexpr int D = 64;                    // head_dim of the tiles (smaller head_dim: zero-padded by the TMA unit, see attn_tc.cu)
constexpr int BM = 128;                  // query rows per warpgroup (one Q block)
constexpr int BN = 64;                   // keys per tile
constexpr int ST = 4;                    // K / V ring stages
constexpr int Q_BYTES = BM * D * 2;      // 16 KB
constexpr int KV_BYTES = BN * D * 2;     //  8 KB
constexpr int kThreads = 384;
constexpr int kRegsControl = 48, kRegsSoftmax = 96;   // 128 * 48 + 256 * 96 = 30720 = 384 * 80 (the launch allocation)
constexpr uint32_t kTmemCols = 256;
constexpr uint32_t TMEM_S = 0, TMEM_ACC = 128;
constexpr float kRescaleThreshold = 8.f;  // log2 units
constexpr int kPolyPairs = PAID_DW_POLY_PAIRS;
constexpr int kSmemBytes = 1024 + 2 * Q_BYTES + ST * 2 * KV_BYTES + 512;

struct DwArgs {
  int mode, fused, N, S, L, heads, head_dim, begin_frame, end_frame;
  int q_pairs, total_items;
  float scale_log2;
  const float* coef;
  void* out;
  int accumulate;
  float out_scale;
  const float* out_frame_scale;
  int per_frame0, per_frame1;  // slot K/V map has one matrix per frame (1) or a single shared matrix (0)
  int wide;                    // output rows of a head start on 32-byte boundaries: 32-byte stores
};

struct Barriers {
  uint64_t q_full[2], q_empty[2];                              // per warpgroup: its Q block
  uint64_t k_full[ST], k_empty[ST], v_full[ST], v_empty[ST];   // shared ring: every tile is consumed by both issuers
  uint64_t s_full[2], p_full[2];                               // per warpgroup: scores ready / P written
  uint64_t acc_final[2], acc_empty[2];                         // per warpgroup: P.V of the item landed / accumulator drained
  uint32_t tmem_slot;
};
static_assert(sizeof(Barriers) <= 512, "barrier block");

// one work item: two adjacent 128-row Q blocks of one head of one frame and the K/V slots they attend to (in order)
struct Item {
  int n, head, row0;
  int nseg;       // 1 or 2 key segments
  int slot0;      // K/V slot of the first segment (0: the frame's own K/V, 1: the lerped endpoint K/V of INNER); a
                  // second segment is always slot 1
  float w;        // output weight (1 for both modes; kept for symmetry with FramePlan)
};

__device__ __forceinline__ int frame_of_order(int z, int N) {
  // interior frames first, the two (cheaper in INNER mode) endpoint frames last
  if (N < 3) return z;
  return z < N - 2 ? z + 1 : (z == N - 2 ? 0 : N - 1);
}

__device__ __forceinline__ Item decode_item(int idx, const DwArgs& a) {
  Item it;
  const int qp = idx % a.q_pairs;
  const int r = idx / a.q_pairs;
  it.head = r % a.heads;
  it.n = frame_of_order(r / a.heads, a.N);
  it.row0 = qp * 2 * BM;
  const float c = a.mode == PAID_PLAIN ? 0.f : a.coef[it.n];
  const FramePlan p = make_frame_plan(a.mode, a.fused, it.n, a.begin_frame, a.end_frame, c);
  it.nseg = (p.use0 ? 1 : 0) + (p.use1 ? 1 : 0);
  it.slot0 = p.use0 ? 0 : 1;
  it.w = p.wA;
  return it;
}

// element pair p (0..15) of a 32-column half goes to the FMA pipe: kPolyPairs of 16, evenly spread
__host__ __device__ constexpr bool pair_on_fma_pipe(int p) { return ptx::pair_on_fma_pipe(p, kPolyPairs); }

// One 64-key score tile of one query row: P = 2^(scale_log2 * (s - m_ref)) written over S as packed 16-bit, row sum
// added to l.  `first`: the accumulator of this warpgroup is still empty, m_ref becomes the tile's true maximum.
template <typename T>
__device__ __forceinline__ void softmax_tile(uint32_t s_addr, uint32_t acc_addr, int valid, bool first, float sl2,
                                             float& m_ref, float& l, long long* tr) {
  if (first) {
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (h * 32 >= valid) break;
      uint32_t sr[32];
      ptx::tmem_ld32(s_addr + h * 32, sr);
      ptx::tmem_wait_ld();
#pragma unroll
      for (int e = 0; e < 32; e += 2) {
        mx0 = fmaxf(mx0, h * 32 + e < valid ? __uint_as_float(sr[e]) : -INFINITY);
        mx1 = fmaxf(mx1, h * 32 + e + 1 < valid ? __uint_as_float(sr[e + 1]) : -INFINITY);
      }
    }
    m_ref = fmaxf(mx0, mx1);
  }
  uint32_t pk[32];
  float sum;
#pragma unroll 1
  for (;;) {
    const float neg = -m_ref * sl2;
    const float2 sl2v = make_float2(sl2, sl2), negv = make_float2(neg, neg);
    float mx0 = -INFINITY, mx1 = -INFINITY;
    float2 sumA = make_float2(0.f, 0.f), sumB = make_float2(0.f, 0.f);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (h * 32 >= valid) {  // ragged last tile: no key in this half (CTA-uniform branch): P = 0, nothing to exponentiate
#pragma unroll
        for (int e = 0; e < 16; ++e) pk[h * 16 + e] = 0u;
        continue;
      }
      uint32_t sr[32];
      TR_T(p0);
      ptx::tmem_ld32(s_addr + h * 32, sr);
      ptx::tmem_wait_ld();
      TR_T(p1);
      if (valid < h * 32 + 32) {  // ragged half (kept a real branch by the asm statement)
        asm volatile("" ::: "memory");
#pragma unroll
        for (int e = 0; e < 32; ++e)
          if (h * 32 + e >= valid) sr[e] = __float_as_uint(-INFINITY);
      }
      if (!first) {
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          mx0 = fmaxf(mx0, __uint_as_float(sr[e]));
          mx1 = fmaxf(mx1, __uint_as_float(sr[e + 1]));
        }
      }
#pragma unroll
      for (int e = 0; e < 32; e += 2) {
        float2 x = ptx::fma2(make_float2(__uint_as_float(sr[e]), __uint_as_float(sr[e + 1])), sl2v, negv);
        if (pair_on_fma_pipe(e / 2)) x = ptx::exp2_poly2(x);
        sr[e] = __float_as_uint(x.x); sr[e + 1] = __float_as_uint(x.y);
      }
      TR_T(p2);
#pragma unroll
      for (int e = 0; e < 32; ++e)
        if (!pair_on_fma_pipe(e / 2)) sr[e] = __float_as_uint(ptx::ex2v(__uint_as_float(sr[e])));
      TR_T(p3);
#pragma unroll
      for (int e = 0; e < 32; e += 4) {
        const float2 x0 = make_float2(__uint_as_float(sr[e]), __uint_as_float(sr[e + 1]));
        const float2 x1 = make_float2(__uint_as_float(sr[e + 2]), __uint_as_float(sr[e + 3]));
        sumA = ptx::add2(sumA, x0);
        sumB = ptx::add2(sumB, x1);
        pk[h * 16 + e / 2] = pack2<T>(x0.x, x0.y);
        pk[h * 16 + e / 2 + 1] = pack2<T>(x1.x, x1.y);
      }
      TR_T(p4);
#ifdef PAID_DW_TRACE
      tr[0] += p1 - p0; tr[1] += p2 - p1; tr[2] += p3 - p2; tr[3] += p4 - p3;
#endif
    }
    sum = (sumA.x + sumA.y) + (sumB.x + sumB.y);
    if (first) break;  // the reference IS the tile maximum
    const float mx = fmaxf(mx0, mx1);
    const bool grow = (mx - m_ref) * sl2 > kRescaleThreshold;
    if (!__any_sync(0xffffffffu, grow)) break;
    // rare: raise the reference of the rows that grew, rescale this warpgroup's accumulator, redo the tile.  The P.V
    // products of all earlier tiles of this warpgroup have landed: the commit that completed s_full for THIS tile was
    // issued after them and tracks every earlier tcgen05.mma of the issuing thread.
    const float m_new = grow ? mx : m_ref;
    const float alpha = ptx::ex2((m_ref - m_new) * sl2);
    l *= alpha;
    m_ref = m_new;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      uint32_t o[32];
      ptx::tmem_ld32(acc_addr + h * 32, o);
      ptx::tmem_wait_ld();
#pragma unroll
      for (int e = 0; e < 32; e += 2) {
        const float2 r = ptx::mul2(make_float2(__uint_as_float(o[e]), __uint_as_float(o[e + 1])), make_float2(alpha, alpha));
        o[e] = __float_as_uint(r.x); o[e + 1] = __float_as_uint(r.y);
      }
      ptx::tmem_st32(acc_addr + h * 32, o);
    }
    ptx::tmem_wait_st();
  }
  TR_T(p5);
  ptx::tmem_st32(s_addr, pk);  // P over the low half of this S buffer (S is not needed any more)
  l += sum;
  ptx::tmem_wait_st();
  TR_T(p6);
#ifdef PAID_DW_TRACE
  tr[4] += p6 - p5;
#endif
}

template <typename T>
__global__ void __launch_bounds__(kThreads, 2)
attn_dw_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK0,
               const __grid_constant__ CUtensorMap tmV0, const __grid_constant__ CUtensorMap tmK1,
               const __grid_constant__ CUtensorMap tmV1, const DwArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                         // [2 warpgroups][128][64]
  uint8_t* sK = sQ + 2 * Q_BYTES;             // [ST][64][64]
  uint8_t* sV = sK + ST * KV_BYTES;           // [ST][64][64]
  Barriers* bar = reinterpret_cast<Barriers*>(sV + ST * KV_BYTES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && ptx::elect_one()) {
    ptx::prefetch_tmap(&tmQ);
    ptx::prefetch_tmap(&tmK0); ptx::prefetch_tmap(&tmV0);
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(&bar->q_full[b], 1); ptx::mbar_init(&bar->q_empty[b], 1);   // the issuer's last Q K^T of the item
      ptx::mbar_init(&bar->s_full[b], 1); ptx::mbar_init(&bar->p_full[b], 4);    // one arrive per softmax warp
      ptx::mbar_init(&bar->acc_final[b], 1); ptx::mbar_init(&bar->acc_empty[b], 4);
    }
    for (int s = 0; s < ST; ++s) {
      ptx::mbar_init(&bar->k_full[s], 1); ptx::mbar_init(&bar->k_empty[s], 2);   // both issuers have read the tile
      ptx::mbar_init(&bar->v_full[s], 1); ptx::mbar_init(&bar->v_empty[s], 2);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) { ptx::tmem_alloc(&bar->tmem_slot, kTmemCols); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = bar->tmem_slot;
  ptx::pdl_launch_dependents();  // the next kernel may begin its prologue
  ptx::pdl_wait();               // everything above overlapped the previous kernel's tail; its results are visible now

  const int tiles = (a.L + BN - 1) / BN;

  if (warp < 4) {
    ptx::setmaxnreg_dec<kRegsControl>();
    if (warp == 0) {
      // ================================ TMA producer ================================
      // The whole warp walks the loops (warp-uniform control flow keeps the counters and addresses in uniform
      // registers); one elected lane arms the barriers and issues the TMA loads.  A Q block that starts at or beyond
      // row S (odd number of 128-row blocks) is zero-filled by the TMA unit; its warpgroup computes on zeros and stores nothing.
      int kc = 0, itc = 0;
      for (int idx = blockIdx.x; idx < a.total_items; idx += gridDim.x, ++itc) {
        const Item it = decode_item(idx, a);
        for (int b = 0; b < 2; ++b) {
          ptx::mbar_wait(&bar->q_empty[b], (itc & 1) ^ 1);
          if (ptx::elect_one()) {
            ptx::mbar_arrive_expect_tx(&bar->q_full[b], Q_BYTES);
            ptx::tma_load_4d(sQ + b * Q_BYTES, &tmQ, &bar->q_full[b], 0, it.head, it.row0 + b * BM, it.n);
          }
        }
        for (int g = 0; g < it.nseg; ++g) {
          const int slot = g == 0 ? it.slot0 : 1;
          const CUtensorMap* mk = slot == 0 ? &tmK0 : &tmK1;
          const CUtensorMap* mv = slot == 0 ? &tmV0 : &tmV1;
          const int fr = (slot == 0 ? a.per_frame0 : a.per_frame1) ? it.n : 0;
          for (int i = 0; i < tiles; ++i, ++kc) {
            const int s = kc % ST;
            const uint32_t ph = (kc / ST) & 1;
            ptx::mbar_wait(&bar->k_empty[s], ph ^ 1);
            if (ptx::elect_one()) {
              ptx::mbar_arrive_expect_tx(&bar->k_full[s], KV_BYTES);
              ptx::tma_load_4d(sK + s * KV_BYTES, mk, &bar->k_full[s], 0, it.head, i * BN, fr);
            }
            ptx::mbar_wait(&bar->v_empty[s], ph ^ 1);
            if (ptx::elect_one()) {
              ptx::mbar_arrive_expect_tx(&bar->v_full[s], KV_BYTES);
              ptx::tma_load_4d(sV + s * KV_BYTES, mv, &bar->v_full[s], 0, it.head, i * BN, fr);
            }
          }
        }
      }
    } else if (warp <= 2) {
      // ================================ MMA issuers =================================
      // Warp 1 issues for warpgroup 0, warp 2 for warpgroup 1: S_b = Q_b K_j^T and acc_b += P_b V_j touch disjoint TMEM
      // columns, so the two issue streams need no ordering between them and neither softmax warpgroup waits behind the
      // other's tile.  Both read every K / V stage; a stage is released when both have committed it.  Warp-uniform loops;
      // one elected lane issues tcgen05.mma / commit.
      const int b = warp - 1;
      constexpr int fmt = std::is_same<T, __nv_bfloat16>::value ? 1 : 0;
      constexpr uint32_t idesc_qk = ptx::make_idesc(BM, BN, fmt, 0);  // S = Q K^T : both operands K-major
      constexpr uint32_t idesc_pv = ptx::make_idesc(BM, D, fmt, 1);   // acc += P V : V is N(=d)-contiguous
      const uint32_t q_addr = ptx::smem_u32(sQ) + b * Q_BYTES, k_addr = ptx::smem_u32(sK), v_addr = ptx::smem_u32(sV);
      const uint32_t s_t = tmem + TMEM_S + b * BN, acc_t = tmem + TMEM_ACC + b * D;
      auto issue_qk = [&](int s) {
        const uint64_t qd = ptx::make_smem_desc_sw128(q_addr, 16, 1024);
        const uint64_t kd = ptx::make_smem_desc_sw128(k_addr + s * KV_BYTES, 16, 1024);
#pragma unroll
        for (int k = 0; k < D / 16; ++k) ptx::mma_ss(s_t, qd + 2 * k, kd + 2 * k, idesc_qk, k != 0);
      };
      int kc = 0, itc = 0;
      uint32_t pcnt = 0;   // P tiles consumed so far (p_full[b] phase)
      TR_DECL(tr_v = 0, tr_k = 0, tr_p = 0, tr_iss = 0, tr_q = 0, tr_steps = 0);
      TR_T(tr_begin);
      for (int idx = blockIdx.x; idx < a.total_items; idx += gridDim.x, ++itc) {
        const Item it = decode_item(idx, a);
        const int total_steps = it.nseg * tiles;
        TR_T(tq0);
        ptx::mbar_wait(&bar->q_full[b], itc & 1);
        TR_T(tq1); TR_ADD(tr_q, tq0, tq1);
        {
          // scores of the first step (S_b is free: the P.V products of the previous item were issued before this point)
          const int s = kc % ST;
          ptx::mbar_wait(&bar->k_full[s], (kc / ST) & 1);
          ptx::tc_fence_after();
          if (ptx::elect_one()) {
            issue_qk(s);
            ptx::tc_commit(&bar->s_full[b]);
            ptx::tc_commit(&bar->k_empty[s]);
            if (total_steps == 1) ptx::tc_commit(&bar->q_empty[b]);   // the last Q K^T of the item
          }
        }
        for (int j = 0; j < total_steps; ++j) {
          const int s = (kc + j) % ST, s2 = (kc + j + 1) % ST;
          const bool more = j + 1 < total_steps;
          TR_T(t0);
          ptx::mbar_wait(&bar->v_full[s], ((kc + j) / ST) & 1);
          TR_T(t1);
          if (more) ptx::mbar_wait(&bar->k_full[s2], ((kc + j + 1) / ST) & 1);
          TR_T(t2);
          ptx::mbar_wait(&bar->p_full[b], pcnt & 1);
          ++pcnt;
          TR_T(t3); TR_ADD(tr_v, t0, t1); TR_ADD(tr_k, t1, t2); TR_ADD(tr_p, t2, t3);
          // the first P.V of an item overwrites acc_b: the epilogue of the previous item must have drained it
          if (j == 0 && itc > 0) ptx::mbar_wait(&bar->acc_empty[b], (itc - 1) & 1);
          ptx::tc_fence_after();
          if (ptx::elect_one()) {
#pragma unroll
            for (int k = 0; k < BN / 16; ++k) {
              // 16 keys per MMA: 8 packed columns of P, 16 rows (2048 B) of the V tile
              const uint64_t vd = ptx::make_smem_desc_sw128(v_addr + s * KV_BYTES + k * 2048, 16, 1024);
              ptx::mma_ts(acc_t, s_t + k * 8, vd, idesc_pv, j != 0 || k != 0);
            }
            if (!more) ptx::tc_commit(&bar->acc_final[b]);   // the P.V products of the item have landed
            if (more) {                                      // S_b is free again (in order after the P.V above)
              issue_qk(s2);
              ptx::tc_commit(&bar->s_full[b]);               // scores of step j + 1
              if (j + 2 >= total_steps) ptx::tc_commit(&bar->q_empty[b]);
            }
            ptx::tc_commit(&bar->v_empty[s]);
            if (more) ptx::tc_commit(&bar->k_empty[s2]);
          }
          TR_T(t4); TR_ADD(tr_iss, t3, t4);
#ifdef PAID_DW_TRACE
          ++tr_steps;
#endif
        }
        kc += total_steps;
      }
#ifdef PAID_DW_TRACE
      if (blockIdx.x == 0 && tr_steps > 0 && ptx::elect_one())
        printf("issuer %d: steps %lld total %lld | per step: wait_v %lld wait_k %lld wait_p %lld issue %lld | wait_q total %lld\n", b,
               tr_steps, clock64() - tr_begin, tr_v / tr_steps, tr_k / tr_steps, tr_p / tr_steps, tr_iss / tr_steps, tr_q);
#endif
    }
  } else {
    ptx::setmaxnreg_inc<kRegsSoftmax>();
    // ================================ softmax warpgroups ==========================
    const int b = (warp - 4) >> 2;   // warpgroup = Q block of the item
    const int quad = warp & 3;       // TMEM lane quadrant of this warp
    const int r = quad * 32 + lane;  // query row within the Q block
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    const uint32_t s_addr = tmem + lane_base + TMEM_S + b * BN;
    const uint32_t acc_addr = tmem + lane_base + TMEM_ACC + b * D;
    const float sl2 = a.scale_log2;
    const int C = a.heads * a.head_dim;
    uint32_t scnt = 0;   // score tiles consumed so far by this warpgroup (s_full phase)
    int itc = 0;
    TR_DECL(tr_s = 0, tr_tile = 0, tr_arr = 0, tr_epi = 0, tr_tiles = 0, tr_ph[5] = {0, 0, 0, 0, 0});
    TR_T(tr_begin);
#ifdef PAID_DW_TRACE
    long long* trp = tr_ph;
#else
    long long* trp = nullptr;
#endif
    for (int idx = blockIdx.x; idx < a.total_items; idx += gridDim.x, ++itc) {
      const Item it = decode_item(idx, a);
      const int total_steps = it.nseg * tiles;
      float m_ref = -INFINITY, l = 0.f;
      int i = 0;                      // tile index inside the current segment
      for (int j = 0; j < total_steps; ++j, ++i) {
        if (i == tiles) i = 0;        // next segment
        TR_T(t0);
        ptx::mbar_wait(&bar->s_full[b], scnt & 1);
        ++scnt;
        ptx::tc_fence_after();
        TR_T(t1);
        softmax_tile<T>(s_addr, acc_addr, a.L - i * BN, j == 0, sl2, m_ref, l, trp);
        TR_T(t2);
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&bar->p_full[b]);
        TR_T(t3); TR_ADD(tr_s, t0, t1); TR_ADD(tr_tile, t1, t2); TR_ADD(tr_arr, t2, t3);
#ifdef PAID_DW_TRACE
        ++tr_tiles;
#endif
      }
      TR_T(te0);
      // ---- epilogue: out = [out +] os * acc / l, the 64 channels of the head for this thread's query row ----
      const float os = a.out_scale * (a.out_frame_scale ? a.out_frame_scale[it.n] : 1.f) * it.w;
      const float inv = os / l;
      const int row = it.row0 + b * BM + r;
      T* dst = (T*)a.out + ((long long)it.n * a.S + row) * C + it.head * a.head_dim;
      ptx::mbar_wait(&bar->acc_final[b], itc & 1);
      ptx::tc_fence_after();
      uint32_t t0[32], t1[32];
      ptx::tmem_ld32(acc_addr, t0);
      ptx::tmem_ld32(acc_addr + 32, t1);
      ptx::tmem_wait_ld();
      // the accumulator is in registers: the issuer may overwrite it with the next item's first P.V
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&bar->acc_empty[b]);
      if (row < a.S) {
#pragma unroll
        for (int v = 0; v < 4; ++v) {       // 16 channels per step
          if (v * 16 >= a.head_dim) break;  // padded head_dim: columns head_dim..63 are zero and are not stored
          const uint32_t* t = v < 2 ? &t0[v * 16] : &t1[(v - 2) * 16];
          const bool both = v * 16 + 8 < a.head_dim;   // head_dim % 16 == 8: the last step has 8 channels
          float o[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) o[e] = inv * __uint_as_float(t[e]);
          if (a.accumulate) {               // out += ...: the IP-Adapter second attention (CTA-uniform branch)
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              if (q == 1 && !both) break;
              const uint4 old = *reinterpret_cast<const uint4*>(dst + v * 16 + q * 8);
              const T* o8 = reinterpret_cast<const T*>(&old);
#pragma unroll
              for (int e = 0; e < 8; ++e) o[q * 8 + e] += to_f32(o8[e]);
            }
          }
          uint32_t w[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) w[e] = pack2<T>(o[2 * e], o[2 * e + 1]);
          if (a.wide && both) {
            ptx::st_global_256(dst + v * 16, w);
          } else {
            *reinterpret_cast<uint4*>(dst + v * 16) = make_uint4(w[0], w[1], w[2], w[3]);
            if (both) *reinterpret_cast<uint4*>(dst + v * 16 + 8) = make_uint4(w[4], w[5], w[6], w[7]);
          }
        }
      }
      TR_T(te1); TR_ADD(tr_epi, te0, te1);
    }
#ifdef PAID_DW_TRACE
    if (blockIdx.x == 0 && lane == 0 && quad == 0 && tr_tiles > 0)
      printf("softmax wg %d: tiles %lld items %d total %lld | per tile: wait_s %lld tile %lld arrive %lld | epilogue per item %lld | "
             "tile phases (both halves): ld %lld max+fma %lld ex2 %lld pack %lld st %lld\n", b,
             tr_tiles, itc, clock64() - tr_begin, tr_s / tr_tiles, tr_tile / tr_tiles, tr_arr / tr_tiles, tr_epi / (itc ? itc : 1),
             tr_ph[0] / tr_tiles, tr_ph[1] / tr_tiles, tr_ph[2] / tr_tiles, tr_ph[3] / tr_tiles, tr_ph[4] / tr_tiles);
#endif
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) { __syncwarp(); ptx::tmem_dealloc(tmem, kTmemCols); }
}

template <typename T>
int launch_t(const CUtensorMap* maps, const DwArgs& da, cudaStream_t stream) {
  auto kern = attn_dw_kernel<T>;
  static_assert(2 * (kSmemBytes + 1024) <= 233472, "two CTAs per SM");
  int num_sms = 0;
  PAID_CUDA_CHECK(ensure_kernel_configured((const void*)kern, kSmemBytes, &num_sms));
  const int slots = 2 * (num_sms > 0 ? num_sms : 148);
  dim3 grid(da.total_items < slots ? da.total_items : slots);
  PAID_CUDA_CHECK(launch_pdl(kern, grid, dim3(kThreads), kSmemBytes, stream, maps[0], maps[1], maps[2], maps[3], maps[4], da));
  PAID_LAUNCH_CHECK("attn_dw_kernel");
  return PAID_OK;
}

}  // namespace

bool attn_dw_supported(const CoreArgs& a) {
  static const bool disabled = [] { const char* e = getenv("PAID_ATTN_DW"); return e && e[0] == '0'; }();
  if (disabled) return false;
  if (a.mode != PAID_PLAIN && a.mode != PAID_INNER) return false;
  if (a.head_dim > D || a.head_dim < 16 || a.head_dim % 8) return false;
  const long long items = (long long)a.N * a.heads * ((a.S + 2 * BM - 1) / (2 * BM));
  return items < (1ll << 30) &&
         !(((uintptr_t)a.q | (uintptr_t)a.k | (uintptr_t)a.v | (uintptr_t)a.out | (uintptr_t)a.k1 | (uintptr_t)a.v1) & 15);
}

int launch_attn_dw(const CoreArgs& a, cudaStream_t stream) {
  CUtensorMap maps[5];
  const int hd = a.head_dim;  // the box of every map is D = 64 wide; the TMA unit zero-fills columns >= hd
  const long long C = (long long)a.heads * hd;
  int st = make_tmap_heads(&maps[0], a.q, a.dtype, a.N, a.S, a.heads, hd, (long long)a.S * C, BM);
  // a driver that rejects a box wider than the tensor: nothing was launched, the caller falls back (core_dispatch)
  if (st != PAID_OK) return hd % D ? PAID_EUNSUPPORTED : st;
  const long long kv_frames = a.stride0 ? a.N : 1;
  if ((st = make_tmap_heads(&maps[1], a.k, a.dtype, kv_frames, a.L, a.heads, hd, a.stride0, BN)) != PAID_OK) return st;
  if ((st = make_tmap_heads(&maps[2], a.v, a.dtype, kv_frames, a.L, a.heads, hd, a.stride0, BN)) != PAID_OK) return st;
  DwArgs da{};
  da.per_frame0 = a.stride0 ? 1 : 0;
  if (a.k1) {
    const long long frames = a.stride1 ? a.N : 1;
    if ((st = make_tmap_heads(&maps[3], a.k1, a.dtype, frames, a.L, a.heads, hd, a.stride1, BN)) != PAID_OK) return st;
    if ((st = make_tmap_heads(&maps[4], a.v1, a.dtype, frames, a.L, a.heads, hd, a.stride1, BN)) != PAID_OK) return st;
    da.per_frame1 = a.stride1 ? 1 : 0;
  } else {  // unused slot: any valid descriptor
    maps[3] = maps[1];
    maps[4] = maps[2];
    da.per_frame1 = 1;
  }
  da.mode = a.mode; da.fused = a.fused; da.N = a.N; da.S = a.S; da.L = a.L; da.heads = a.heads; da.head_dim = hd;
  da.begin_frame = a.begin_frame; da.end_frame = a.end_frame;
  da.q_pairs = (a.S + 2 * BM - 1) / (2 * BM);
  da.total_items = a.N * a.heads * da.q_pairs;
  da.scale_log2 = a.scale * kLog2e;
  da.coef = a.coef; da.out = a.out;
  da.accumulate = a.accumulate; da.out_scale = a.out_scale; da.out_frame_scale = a.out_frame_scale;
  da.wide = hd % 16 == 0 && !((uintptr_t)a.out & 31);
  return a.dtype == PAID_F16 ? launch_t<__half>(maps, da, stream) : launch_t<__nv_bfloat16>(maps, da, stream);
}

}  // namespace paid
