// C ABI of libpaid_attn.so (include/paid_attn.h): validation, workspace carving, kernel dispatch.
#include <cstring>
#include <mutex>
#include <utility>
#include <vector>

#include "paid_common.cuh"

namespace paid {

char* error_buffer() {
  static thread_local char buf[512] = {0};
  return buf;
}
const char** last_kernel_slot() {
  static thread_local const char* k = "";
  return &k;
}
int fail(int status, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(error_buffer(), 512, fmt, ap);
  va_end(ap);
  return status;
}
std::atomic<uint64_t>& launch_counter() {
  static std::atomic<uint64_t> c{0};
  return c;
}

cudaError_t ensure_kernel_configured(const void* kernel, int smem_bytes, int* num_sms) {
  struct Entry { const void* kernel; int dev; };
  static std::mutex mu;
  static std::vector<Entry> done;
  static int sms[64] = {0};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lk(mu);
  bool found = false;
  for (const Entry& en : done) found = found || (en.kernel == kernel && en.dev == dev);
  if (!found) {
    if ((e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes)) != cudaSuccess) return e;
    done.push_back({kernel, dev});
  }
  if (num_sms) {
    if (dev < 0 || dev >= 64) return cudaDeviceGetAttribute(num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms[dev] == 0 && (e = cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
    *num_sms = sms[dev];
  }
  return cudaSuccess;
}

static inline uint64_t align256(uint64_t b) { return (b + 255) & ~uint64_t(255); }

struct Workspace {
  uint64_t q, k, v, h, kx, vx, total;
};

static Workspace plan_workspace(const PaidAttnParams* p) {
  Workspace w{};
  const uint64_t es = 2;
  const uint64_t frames = (uint64_t)p->N + (uint64_t)p->plain_tail;
  uint64_t nsc = align256(frames * p->S * p->C * es);
  uint64_t nlc = align256(frames * p->L * p->C * es);
  uint64_t off = 0;
  w.q = off; off += nsc;
  w.k = off; off += nlc;
  w.v = off; off += nlc;
  w.h = off; off += nsc;
  if (p->mode == PAID_INNER) { w.kx = off; off += nlc; w.vx = off; off += nlc; }
  w.total = off;
  return w;
}

static int validate(const PaidAttnParams* p, bool need_io) {
  if (!p) return fail(PAID_EINVAL, "params is NULL");
  if (p->struct_size != sizeof(PaidAttnParams))
    return fail(PAID_EINVAL, "PaidAttnParams.struct_size %u != %zu (ABI mismatch)", p->struct_size, sizeof(PaidAttnParams));
  if (p->dtype != PAID_F16 && p->dtype != PAID_BF16) return fail(PAID_EINVAL, "dtype must be PAID_F16 or PAID_BF16");
  if (p->mode < PAID_PLAIN || p->mode > PAID_INNER) return fail(PAID_EINVAL, "bad mode %d", p->mode);
  if (p->N <= 0 || p->S <= 0 || p->L <= 0 || p->C <= 0 || p->Cc <= 0 || p->heads <= 0)
    return fail(PAID_EINVAL, "sizes must be positive (N=%d S=%d L=%d C=%d Cc=%d heads=%d)", p->N, p->S, p->L, p->C,
                p->Cc, p->heads);
  if (p->plain_tail < 0) return fail(PAID_EINVAL, "plain_tail=%d must not be negative", p->plain_tail);
  if (p->C % p->heads) return fail(PAID_EINVAL, "C=%d is not a multiple of heads=%d", p->C, p->heads);
  if (p->C % 8 || p->Cc % 8) return fail(PAID_EUNSUPPORTED, "C and Cc must be multiples of 8 (16-byte rows)");
  if (!p->ctx && (p->L != p->S || p->Cc != p->C))
    return fail(PAID_EINVAL, "self-attention (ctx == NULL) needs L == S and Cc == C");
  if (!need_io) return PAID_OK;
  if (!p->x || !p->wq || !p->wo || !p->y) return fail(PAID_EINVAL, "x, wq, wo, y must be non-NULL");
  if ((p->k_pre == nullptr) != (p->v_pre == nullptr)) return fail(PAID_EINVAL, "k_pre and v_pre must be given together");
  if (!p->k_pre && (!p->wk || !p->wv)) return fail(PAID_EINVAL, "wk, wv must be non-NULL (or k_pre / v_pre given)");
  if (p->kv_pre_broadcast && (!p->k_pre || p->mode != PAID_PLAIN))
    return fail(PAID_EINVAL, "kv_pre_broadcast needs k_pre / v_pre and PLAIN mode");
  if (p->plain_tail && p->kv_pre_broadcast)
    return fail(PAID_EINVAL, "plain_tail needs per-frame k_pre / v_pre (kv_pre_broadcast must be 0)");
  if (p->mode != PAID_PLAIN) {
    if (!p->coef) return fail(PAID_EINVAL, "coef is NULL");
    if (!p->kv_ext && (p->begin_frame < 0 || p->begin_frame >= p->N || p->end_frame < 0 || p->end_frame >= p->N))
      return fail(PAID_EINVAL, "begin_frame/end_frame (%d,%d) must index this batch when kv_ext is NULL", p->begin_frame,
                  p->end_frame);
    if (p->begin_frame >= p->N || p->end_frame >= p->N) return fail(PAID_EINVAL, "endpoint frame index out of range");
  }
  return PAID_OK;
}

static int linear(const void* x, const void* w, const void* bias, void* y, long long M, int Nout, int K, int dtype,
                  uint32_t flags, cudaStream_t stream) {
  profile_mark_begin(stream, PAID_PROFILE_LINEAR, M, Nout, K, 1, 2.0 * M * Nout * K);
  const int st = (!(flags & PAID_FLAG_GENERIC_KERNELS) && linear_tc_supported(M, Nout, K))
                     ? launch_linear_tc(x, w, bias, y, M, Nout, K, dtype, stream)
                     : launch_linear_generic(x, w, bias, y, M, Nout, K, dtype, stream);
  profile_mark_end(stream);
  return st;
}

// up to three projections of the same input in one launch (q/k/v of self-attention, k/v of cross-attention)
static int linear_grouped(const void* x, const void* const* w, void* const* y, int groups, long long M, int Nout, int K,
                          int dtype, uint32_t flags, cudaStream_t stream) {
  profile_mark_begin(stream, PAID_PROFILE_LINEAR, M, Nout, K, groups, 2.0 * M * Nout * K * groups);
  int st = PAID_OK;
  if (!(flags & PAID_FLAG_GENERIC_KERNELS) && linear_tc_supported(M, Nout, K)) {
    st = launch_linear_tc_grouped(x, w, nullptr, y, groups, M, Nout, K, dtype, stream);
  } else {
    for (int g = 0; g < groups && st == PAID_OK; ++g) st = launch_linear_generic(x, w[g], nullptr, y[g], M, Nout, K, dtype, stream);
  }
  profile_mark_end(stream);
  return st;
}

// ---- measurement hook: CUDA events around the attention-core and GEMM kernels ----------------------
struct ProfileRecord {
  cudaEvent_t begin = nullptr, end = nullptr;
  int kind = 0;            // PAID_PROFILE_ATTENTION / _LINEAR / _LINEAR_GEGLU
  long long d[4] = {0, 0, 0, 0};
  double flops = 0;
};
struct ProfileState {
  std::mutex mu;
  bool on = false;
  std::vector<ProfileRecord> used;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pool;
  ProfileRecord current;
};
static ProfileState& prof() {
  static ProfileState p;
  return p;
}
void profile_mark_begin(cudaStream_t stream, int kind, long long d0, long long d1, long long d2, long long d3, double flops) {
  ProfileState& ps = prof();
  if (!ps.on) return;
  std::lock_guard<std::mutex> lk(ps.mu);
  ProfileRecord r;
  if (!ps.pool.empty()) { r.begin = ps.pool.back().first; r.end = ps.pool.back().second; ps.pool.pop_back(); }
  else if (cudaEventCreate(&r.begin) != cudaSuccess || cudaEventCreate(&r.end) != cudaSuccess) {
    ps.current = ProfileRecord{};
    return;
  }
  r.kind = kind; r.d[0] = d0; r.d[1] = d1; r.d[2] = d2; r.d[3] = d3; r.flops = flops;
  ps.current = r;
  cudaEventRecord(r.begin, stream);
}
void profile_mark_end(cudaStream_t stream) {
  ProfileState& ps = prof();
  if (!ps.on || !ps.current.begin) return;
  std::lock_guard<std::mutex> lk(ps.mu);
  cudaEventRecord(ps.current.end, stream);
  ps.used.push_back(ps.current);
  ps.current = ProfileRecord{};
}
static double algorithmic_flops(const CoreArgs& a) {
  const double A = 2.0 * a.N * a.S * a.L * a.heads * a.head_dim;
  if (a.mode == PAID_OUTER) return (a.fused ? 6.0 : 4.0) * A;
  if (a.mode == PAID_INNER) return (a.fused ? 4.0 : 2.0) * A;
  return 2.0 * A;
}

static int core_dispatch(const CoreArgs& a, uint32_t flags, cudaStream_t stream) {
  int st;
  profile_mark_begin(stream, PAID_PROFILE_ATTENTION, a.S, a.L, (long long)a.heads * a.head_dim, 16 * a.mode + a.fused,
                     algorithmic_flops(a));
  static std::atomic<bool> pad_rejected{false};  // a zero-padded head_dim needs a TMA box wider than the tensor
  const bool padded = a.head_dim % 64 != 0;
  if (!(flags & PAID_FLAG_GENERIC_KERNELS) && attn_tc_supported(a) && !(padded && pad_rejected.load())) {
    *last_kernel_slot() = padded ? "tcgen05-padded" : "tcgen05";
    // single-stream modes at head_dim <= 64: the persistent dual-warpgroup kernel; OUTER and wide heads: attn_tc.cu
    st = (!(flags & PAID_FLAG_ONE_WARPGROUP) && attn_dw_supported(a)) ? launch_attn_dw(a, stream) : launch_attn_tc(a, stream);
    if (st == PAID_EUNSUPPORTED && padded) {  // descriptor refused before any launch: other CUDA kernel family
      pad_rejected.store(true);
      *last_kernel_slot() = "generic";
      st = launch_attn_generic(a, stream);
    }
  } else {
    *last_kernel_slot() = "generic";
    st = launch_attn_generic(a, stream);
  }
  profile_mark_end(stream);
  return st;
}

// resolve slot 1 / slot 2 sources; for INNER run the endpoint lerp into (kx, vx)
static int resolve_slots(CoreArgs& a, const void* kv_ext, void* kx, void* vx, cudaStream_t stream) {
  const long long LC = (long long)a.L * a.heads * a.head_dim;
  const char* kb; const char* vb; const char* ke; const char* ve;
  if (kv_ext) {
    const char* e = (const char*)kv_ext;
    kb = e; vb = e + LC * 2; ke = e + 2 * LC * 2; ve = e + 3 * LC * 2;
  } else {
    kb = (const char*)a.k + (long long)a.begin_frame * LC * 2;
    vb = (const char*)a.v + (long long)a.begin_frame * LC * 2;
    ke = (const char*)a.k + (long long)a.end_frame * LC * 2;
    ve = (const char*)a.v + (long long)a.end_frame * LC * 2;
  }
  a.k1 = a.v1 = a.k2 = a.v2 = nullptr;
  a.stride1 = a.stride2 = 0;
  if (a.mode == PAID_OUTER) {
    a.k1 = kb; a.v1 = vb; a.k2 = ke; a.v2 = ve;
  } else if (a.mode == PAID_INNER) {
    int st = launch_lerp_endpoints(kb, vb, ke, ve, a.coef, kx, vx, a.N, LC, a.dtype, stream);
    if (st != PAID_OK) return st;
    a.k1 = kx; a.v1 = vx; a.stride1 = LC;
  }
  return PAID_OK;
}

}  // namespace paid

using namespace paid;

extern "C" {

int paid_attn_abi_version(void) { return PAID_ABI_VERSION; }

const char* paid_attn_last_error(void) { return error_buffer(); }

uint64_t paid_attn_launch_count(void) { return launch_counter().load(std::memory_order_relaxed); }

const char* paid_attn_last_kernel(void) { return *last_kernel_slot(); }

int paid_attn_profile_enable(int on) {
  ProfileState& ps = prof();
  std::lock_guard<std::mutex> lk(ps.mu);
  ps.on = on != 0;
  return PAID_OK;
}

static int profile_collect(std::vector<PaidProfileRow>& rows) {
  ProfileState& ps = prof();
  for (auto& r : ps.used) {
    PAID_CUDA_CHECK(cudaEventSynchronize(r.end));
    float t = 0;
    PAID_CUDA_CHECK(cudaEventElapsedTime(&t, r.begin, r.end));
    PaidProfileRow* row = nullptr;
    for (auto& q : rows)
      if (q.kind == r.kind && q.d[0] == r.d[0] && q.d[1] == r.d[1] && q.d[2] == r.d[2] && q.d[3] == r.d[3]) { row = &q; break; }
    if (!row) {
      PaidProfileRow q{};
      q.kind = r.kind;
      for (int i = 0; i < 4; ++i) q.d[i] = r.d[i];
      rows.push_back(q);
      row = &rows.back();
    }
    row->launches += 1; row->total_ms += t; row->flops += r.flops;
  }
  return PAID_OK;
}
static void profile_reset() {
  ProfileState& ps = prof();
  for (auto& r : ps.used) ps.pool.push_back({r.begin, r.end});
  ps.used.clear();
}

int paid_attn_profile_read(double* total_ms, uint64_t* launches, double* alg_flops, int reset) {
  ProfileState& ps = prof();
  std::lock_guard<std::mutex> lk(ps.mu);
  std::vector<PaidProfileRow> rows;
  int st = profile_collect(rows);
  if (st != PAID_OK) return st;
  double ms = 0, fl = 0;
  uint64_t n = 0;
  for (auto& q : rows)
    if (q.kind == PAID_PROFILE_ATTENTION) { ms += q.total_ms; n += q.launches; fl += q.flops; }
  if (total_ms) *total_ms = ms;
  if (launches) *launches = n;
  if (alg_flops) *alg_flops = fl;
  if (reset) profile_reset();
  return PAID_OK;
}

int paid_attn_profile_rows(PaidProfileRow* out, uint64_t capacity, uint64_t* count, int reset) {
  if (!count) return fail(PAID_EINVAL, "paid_attn_profile_rows: count is NULL");
  ProfileState& ps = prof();
  std::lock_guard<std::mutex> lk(ps.mu);
  std::vector<PaidProfileRow> rows;
  int st = profile_collect(rows);
  if (st != PAID_OK) return st;
  *count = rows.size();
  if (out)
    for (uint64_t i = 0; i < rows.size() && i < capacity; ++i) out[i] = rows[i];
  if (reset && (out || capacity == 0)) profile_reset();
  return PAID_OK;
}

uint64_t paid_attn_workspace_bytes(const PaidAttnParams* p) {
  if (validate(p, false) != PAID_OK) return 0;
  return plan_workspace(p).total;
}

uint64_t paid_attn_core_workspace_bytes(const PaidCoreParams* p) {
  if (!p || p->struct_size != sizeof(PaidCoreParams)) return 0;
  if (p->mode != PAID_INNER) return 0;
  return 2 * align256((uint64_t)p->N * p->L * p->heads * p->head_dim * 2);
}

int paid_linear(const void* x, const void* w, const void* bias, void* y, int64_t M, int32_t Nout, int32_t K,
                int32_t dtype, uint32_t flags, void* cuda_stream) {
  if (!x || !w || !y) return fail(PAID_EINVAL, "paid_linear: x, w, y must be non-NULL");
  if (M <= 0 || Nout <= 0 || K <= 0) return fail(PAID_EINVAL, "paid_linear: sizes must be positive");
  if (dtype != PAID_F16 && dtype != PAID_BF16) return fail(PAID_EINVAL, "paid_linear: bad dtype");
  return linear(x, w, bias, y, M, Nout, K, dtype, flags, (cudaStream_t)cuda_stream);
}

int paid_linear_geglu(const void* x, const void* w, const void* bias, void* y, int64_t M, int32_t D, int32_t K,
                      int32_t dtype, uint32_t flags, void* cuda_stream) {
  if (!x || !w || !y) return fail(PAID_EINVAL, "paid_linear_geglu: x, w, y must be non-NULL");
  if (M <= 0 || D <= 0 || K <= 0) return fail(PAID_EINVAL, "paid_linear_geglu: sizes must be positive");
  if (dtype != PAID_F16 && dtype != PAID_BF16) return fail(PAID_EINVAL, "paid_linear_geglu: bad dtype");
  cudaStream_t stream = (cudaStream_t)cuda_stream;
  profile_mark_begin(stream, PAID_PROFILE_LINEAR_GEGLU, M, D, K, 1, 4.0 * M * D * K);
  const int st = (!(flags & PAID_FLAG_GENERIC_KERNELS) && linear_geglu_tc_supported(M, D, K))
                     ? launch_linear_geglu_tc(x, w, bias, y, M, D, K, dtype, stream)
                     : launch_linear_geglu_generic(x, w, bias, y, M, D, K, dtype, stream);
  profile_mark_end(stream);
  return st;
}

int paid_geglu(const void* h, void* out, int64_t M, int32_t D, int32_t dtype, void* cuda_stream) {
  if (!h || !out) return fail(PAID_EINVAL, "paid_geglu: h and out must be non-NULL");
  if (M <= 0 || D <= 0 || D % 8) return fail(PAID_EINVAL, "paid_geglu: M > 0 and D a positive multiple of 8");
  if (dtype != PAID_F16 && dtype != PAID_BF16) return fail(PAID_EINVAL, "paid_geglu: bad dtype");
  if (((uintptr_t)h | (uintptr_t)out) & 15) return fail(PAID_EINVAL, "paid_geglu: pointers must be 16-byte aligned");
  return launch_geglu(h, out, M, D, dtype, (cudaStream_t)cuda_stream);
}

int paid_add_layer_norm(const void* x, const void* delta, const void* gamma, const void* beta, void* x_out, void* h_out,
                        int64_t rows, int32_t C, float eps, int32_t dtype, void* cuda_stream) {
  if (!x || !gamma || !beta || !h_out) return fail(PAID_EINVAL, "paid_add_layer_norm: x, gamma, beta, h_out must be non-NULL");
  if (delta && !x_out) return fail(PAID_EINVAL, "paid_add_layer_norm: x_out must be given with delta");
  if (rows <= 0 || C <= 0) return fail(PAID_EINVAL, "paid_add_layer_norm: sizes must be positive");
  if (dtype != PAID_F16 && dtype != PAID_BF16) return fail(PAID_EINVAL, "paid_add_layer_norm: bad dtype");
  if (!add_layer_norm_supported(C)) return fail(PAID_EUNSUPPORTED, "paid_add_layer_norm: C=%d must be a multiple of 8, <= 2048", C);
  if (((uintptr_t)x | (uintptr_t)delta | (uintptr_t)gamma | (uintptr_t)beta | (uintptr_t)x_out | (uintptr_t)h_out) & 15)
    return fail(PAID_EINVAL, "paid_add_layer_norm: pointers must be 16-byte aligned");
  return launch_add_layer_norm(x, delta, gamma, beta, x_out, h_out, rows, C, eps, dtype, (cudaStream_t)cuda_stream);
}

int paid_residual_bias_add(const void* a, const void* b, const void* bias, void* out, int64_t rows, int32_t C,
                           int32_t dtype, void* cuda_stream) {
  if (!a || !b || !bias || !out) return fail(PAID_EINVAL, "paid_residual_bias_add: a, b, bias, out must be non-NULL");
  if (rows <= 0 || C <= 0 || C % 8) return fail(PAID_EINVAL, "paid_residual_bias_add: rows > 0 and C a positive multiple of 8");
  if (dtype != PAID_F16 && dtype != PAID_BF16) return fail(PAID_EINVAL, "paid_residual_bias_add: bad dtype");
  if (((uintptr_t)a | (uintptr_t)b | (uintptr_t)bias | (uintptr_t)out) & 15)
    return fail(PAID_EINVAL, "paid_residual_bias_add: pointers must be 16-byte aligned");
  return launch_residual_bias_add(a, b, bias, out, rows, C, dtype, (cudaStream_t)cuda_stream);
}

uint64_t paid_group_norm_workspace_bytes(int32_t N, int64_t HW, int32_t C, int32_t groups) {
  if (!group_norm_supported(C, groups)) return 0;
  return group_norm_workspace_bytes(N, HW, C, groups);
}

int paid_group_norm_nhwc(const void* x, const void* pre_bias, const void* gamma, const void* beta, void* y, void* workspace,
                         uint64_t workspace_bytes, int32_t N, int64_t HW, int32_t C, int32_t groups, float eps,
                         int32_t silu, int32_t dtype, void* cuda_stream) {
  if (!x || !gamma || !beta || !y) return fail(PAID_EINVAL, "paid_group_norm_nhwc: x, gamma, beta, y must be non-NULL");
  if (N <= 0 || N > 65535 || HW <= 0 || C <= 0 || groups <= 0) return fail(PAID_EINVAL, "paid_group_norm_nhwc: bad sizes");
  if (dtype != PAID_F16 && dtype != PAID_BF16) return fail(PAID_EINVAL, "paid_group_norm_nhwc: bad dtype");
  if (!group_norm_supported(C, groups))
    return fail(PAID_EUNSUPPORTED, "paid_group_norm_nhwc: C=%d groups=%d (need C %% 8 == 0, C <= 4096, groups <= 64, C %% groups == 0)", C, groups);
  if (((uintptr_t)x | (uintptr_t)pre_bias | (uintptr_t)gamma | (uintptr_t)beta | (uintptr_t)y | (uintptr_t)workspace) & 15)
    return fail(PAID_EINVAL, "paid_group_norm_nhwc: pointers must be 16-byte aligned");
  const uint64_t need = group_norm_workspace_bytes(N, HW, C, groups);
  if (!workspace || workspace_bytes < need)
    return fail(PAID_EWORKSPACE, "paid_group_norm_nhwc: workspace needs %llu bytes, got %llu", (unsigned long long)need,
                (unsigned long long)workspace_bytes);
  return launch_group_norm_nhwc(x, pre_bias, gamma, beta, y, (float*)workspace, N, HW, C, groups, eps, silu ? 1 : 0, dtype,
                                (cudaStream_t)cuda_stream);
}

int paid_attn_core(const PaidCoreParams* p, void* cuda_stream) {
  if (!p) return fail(PAID_EINVAL, "params is NULL");
  if (p->struct_size != sizeof(PaidCoreParams))
    return fail(PAID_EINVAL, "PaidCoreParams.struct_size %u != %zu (ABI mismatch)", p->struct_size, sizeof(PaidCoreParams));
  if (p->dtype != PAID_F16 && p->dtype != PAID_BF16) return fail(PAID_EINVAL, "bad dtype");
  if (p->mode < PAID_PLAIN || p->mode > PAID_INNER) return fail(PAID_EINVAL, "bad mode %d", p->mode);
  if (p->N <= 0 || p->S <= 0 || p->L <= 0 || p->heads <= 0 || p->head_dim <= 0) return fail(PAID_EINVAL, "sizes must be positive");
  if ((p->heads * p->head_dim) % 8) return fail(PAID_EUNSUPPORTED, "heads*head_dim must be a multiple of 8");
  if (!p->q || !p->k || !p->v || !p->out) return fail(PAID_EINVAL, "q, k, v, out must be non-NULL");
  if (p->mode != PAID_PLAIN) {
    if (!p->coef) return fail(PAID_EINVAL, "coef is NULL");
    if (!p->kv_ext && (p->begin_frame < 0 || p->begin_frame >= p->N || p->end_frame < 0 || p->end_frame >= p->N))
      return fail(PAID_EINVAL, "begin_frame/end_frame must index this batch when kv_ext is NULL");
    if (p->begin_frame >= p->N || p->end_frame >= p->N) return fail(PAID_EINVAL, "endpoint frame index out of range");
  }
  cudaStream_t stream = (cudaStream_t)cuda_stream;
  CoreArgs a{};
  a.dtype = p->dtype; a.mode = p->mode; a.fused = p->fused ? 1 : 0;
  a.N = p->N; a.S = p->S; a.L = p->L; a.heads = p->heads; a.head_dim = p->head_dim;
  a.scale = p->scale; a.begin_frame = p->begin_frame; a.end_frame = p->end_frame;
  a.q = p->q; a.k = p->k; a.v = p->v; a.coef = p->coef; a.out = p->out;
  a.accumulate = p->accumulate ? 1 : 0;
  a.out_scale = p->out_scale == 0.f ? 1.f : p->out_scale;
  a.out_frame_scale = p->out_frame_scale;
  if (p->kv_broadcast && p->mode != PAID_PLAIN) return fail(PAID_EINVAL, "kv_broadcast is only valid in PLAIN mode");
  a.stride0 = p->kv_broadcast ? 0 : (long long)p->L * p->heads * p->head_dim;
  void* kx = nullptr; void* vx = nullptr;
  if (p->mode == PAID_INNER) {
    uint64_t need = paid_attn_core_workspace_bytes(p);
    if (!p->workspace || p->workspace_bytes < need)
      return fail(PAID_EWORKSPACE, "INNER core needs %llu workspace bytes, got %llu", (unsigned long long)need,
                  (unsigned long long)p->workspace_bytes);
    kx = p->workspace;
    vx = (char*)p->workspace + need / 2;
  }
  int st = resolve_slots(a, p->mode == PAID_PLAIN ? nullptr : p->kv_ext, kx, vx, stream);
  if (st != PAID_OK) return st;
  return core_dispatch(a, p->flags, stream);
}

int paid_attn_project_endpoints(const PaidAttnParams* p, int32_t local_frame, void* k_out, void* v_out,
                                void* cuda_stream) {
  int st = validate(p, false);
  if (st != PAID_OK) return st;
  if (!p->x || !p->wk || !p->wv || !k_out || !v_out) return fail(PAID_EINVAL, "x, wk, wv, k_out, v_out must be non-NULL");
  if (local_frame < 0 || local_frame >= p->N) return fail(PAID_EINVAL, "local_frame %d out of range [0,%d)", local_frame, p->N);
  cudaStream_t stream = (cudaStream_t)cuda_stream;
  const char* src = p->ctx ? (const char*)p->ctx : (const char*)p->x;
  src += (long long)local_frame * p->L * p->Cc * 2;
  const void* w[2] = {p->wk, p->wv};
  void* y[2] = {k_out, v_out};
  return linear_grouped(src, w, y, 2, p->L, p->C, p->Cc, p->dtype, p->flags, stream);
}

int paid_attn_project_kv(const PaidAttnParams* p, void* k_out, void* v_out, void* cuda_stream) {
  int st = validate(p, false);
  if (st != PAID_OK) return st;
  if (!p->x || !p->wk || !p->wv || !k_out || !v_out) return fail(PAID_EINVAL, "x, wk, wv, k_out, v_out must be non-NULL");
  const void* src = p->ctx ? p->ctx : p->x;
  const void* w[2] = {p->wk, p->wv};
  void* y[2] = {k_out, v_out};
  return linear_grouped(src, w, y, 2, (long long)p->N * p->L, p->C, p->Cc, p->dtype, p->flags, (cudaStream_t)cuda_stream);
}

int paid_attn_forward(const PaidAttnParams* p, void* cuda_stream) {
  int st = validate(p, true);
  if (st != PAID_OK) return st;
  Workspace ws = plan_workspace(p);
  if (!p->workspace || p->workspace_bytes < ws.total)
    return fail(PAID_EWORKSPACE, "workspace too small: need %llu bytes, got %llu", (unsigned long long)ws.total,
                (unsigned long long)p->workspace_bytes);
  if ((uintptr_t)p->workspace & 255) return fail(PAID_EINVAL, "workspace must be 256-byte aligned");
  cudaStream_t stream = (cudaStream_t)cuda_stream;
  char* base = (char*)p->workspace;
  void* Q = base + ws.q; void* K = base + ws.k; void* V = base + ws.v; void* H = base + ws.h;
  const void* src = p->ctx ? p->ctx : p->x;
  const long long NT = (long long)p->N + p->plain_tail;   // the projections run over the interpolation sequence + CFG rows
  const long long MS = NT * p->S, ML = NT * p->L;

  // interpolation.py:613, 623-624
  if (p->k_pre) {  // K / V of a step-invariant context were projected once per sequence (paid_attn_project_kv)
    K = const_cast<void*>(p->k_pre); V = const_cast<void*>(p->v_pre);
    if ((st = linear(p->x, p->wq, nullptr, Q, MS, p->C, p->C, p->dtype, p->flags, stream)) != PAID_OK) return st;
  } else if (!p->ctx) {  // self-attention: q, k, v share the input -> one launch
    const void* w[3] = {p->wq, p->wk, p->wv};
    void* y[3] = {Q, K, V};
    if ((st = linear_grouped(p->x, w, y, 3, MS, p->C, p->C, p->dtype, p->flags, stream)) != PAID_OK) return st;
  } else {
    if ((st = linear(p->x, p->wq, nullptr, Q, MS, p->C, p->C, p->dtype, p->flags, stream)) != PAID_OK) return st;
    const void* w[2] = {p->wk, p->wv};
    void* y[2] = {K, V};
    if ((st = linear_grouped(src, w, y, 2, ML, p->C, p->Cc, p->dtype, p->flags, stream)) != PAID_OK) return st;
  }

  CoreArgs a{};
  a.dtype = p->dtype; a.mode = p->mode; a.fused = p->fused ? 1 : 0;
  a.N = p->N; a.S = p->S; a.L = p->L; a.heads = p->heads; a.head_dim = p->C / p->heads;
  a.scale = p->scale; a.begin_frame = p->begin_frame; a.end_frame = p->end_frame;
  a.q = Q; a.k = K; a.v = V; a.coef = p->coef; a.out = H;
  a.accumulate = 0; a.out_scale = 1.f; a.out_frame_scale = nullptr;
  a.stride0 = p->kv_pre_broadcast ? 0 : (long long)p->L * p->C;
  void* kx = p->mode == PAID_INNER ? base + ws.kx : nullptr;
  void* vx = p->mode == PAID_INNER ? base + ws.vx : nullptr;
  // the endpoint K/V of a frame-sharded sequence arrive on another stream (NCCL broadcast): wait here, after the local
  // projections have been queued
  if (p->kv_ext_ready_event && p->mode != PAID_PLAIN)
    PAID_CUDA_CHECK(cudaStreamWaitEvent(stream, (cudaEvent_t)p->kv_ext_ready_event, 0));
  if ((st = resolve_slots(a, p->mode == PAID_PLAIN ? nullptr : p->kv_ext, kx, vx, stream)) != PAID_OK) return st;
  // interpolation.py:627-664 / 760-790
  if ((st = core_dispatch(a, p->flags, stream)) != PAID_OK) return st;
  if (p->plain_tail > 0) {  // the unconditional rows of the step: stock attention on frames [N, N + plain_tail)
    const long long qoff = (long long)p->N * p->S * p->C * 2, koff = (long long)p->N * p->L * p->C * 2;
    CoreArgs b{};
    b.dtype = p->dtype; b.mode = PAID_PLAIN; b.fused = 0;
    b.N = p->plain_tail; b.S = p->S; b.L = p->L; b.heads = p->heads; b.head_dim = a.head_dim;
    b.scale = p->scale; b.begin_frame = 0; b.end_frame = p->plain_tail - 1;
    b.q = (const char*)Q + qoff; b.k = (const char*)K + koff; b.v = (const char*)V + koff; b.out = (char*)H + qoff;
    b.accumulate = 0; b.out_scale = 1.f; b.out_frame_scale = nullptr;
    b.stride0 = (long long)p->L * p->C;
    if ((st = core_dispatch(b, p->flags, stream)) != PAID_OK) return st;
  }
  // interpolation.py:666-667
  return linear(H, p->wo, p->bo, p->y, MS, p->C, p->C, p->dtype, p->flags, stream);
}

}  // extern "C"
