/*
 * paid_attn.h -- C ABI of libpaid_attn.so: the PAID / AID interpolated-attention
 * hot path as hand-written sm_100a CUDA.
 *
 * The reference (QY-H00/attention-interpolation-diffusion) is pure Python and has
 * no FFI of its own; the boundary it exposes for this path is the diffusers
 * AttnProcessor protocol.  Each entry point below replaces the body of one
 * reference call and is what a binding written against the reference would bind:
 *
 *   paid_attn_forward            OuterInterpolatedAttnProcessor.__call__  interpolation.py:573-679
 *                                InnerInterpolatedAttnProcessor.__call__  interpolation.py:707-804
 *                                deactivated branch (original_attn)       interpolation.py:581-584, 715-718
 *   paid_attn_core               attn.get_attention_scores + torch.bmm + alpha-lerp
 *                                                                         interpolation.py:627-664, 760-790
 *   paid_linear                  attn.to_q / to_k / to_v / to_out[0]      interpolation.py:613, 623-624, 666
 *   paid_attn_project_kv         attn.to_k / to_v of a step-invariant context, once per sequence  interpolation.py:623-624
 *   paid_attn_project_endpoints  key[0:1], key[-1:], value[0:1], value[-1:] interpolation.py:627-630
 *                                (for frame-sharded execution: the owner rank
 *                                projects, NCCL broadcasts, every rank consumes
 *                                through PaidAttnParams.kv_ext)
 *
 * Conventions
 *   - plain C, no torch types; every pointer is a DEVICE pointer owned by the
 *     caller (PyTorch), including the workspace.  The library allocates no device
 *     memory and keeps no state besides a small host-side TMA-descriptor cache.
 *   - all tensors are dense row-major with the shapes given; 16-byte aligned.
 *   - asynchronous on the given cudaStream_t (passed as void*); no host sync.
 *   - returns PAID_OK (0) or a negative PaidStatus; never throws, never aborts.
 *     paid_attn_last_error() gives a thread-local message for the last failure.
 *   - there is no CPU fallback: without a CUDA device every compute entry point
 *     returns PAID_ECUDA.
 */
#ifndef PAID_ATTN_H_
#define PAID_ATTN_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PAID_ABI_VERSION 3

typedef enum PaidStatus {
  PAID_OK = 0,
  PAID_EINVAL = -1,       /* bad argument (null pointer, size <= 0, C % heads != 0, ...) */
  PAID_EUNSUPPORTED = -2, /* valid but not implemented shape / dtype */
  PAID_ECUDA = -3,        /* CUDA runtime / driver error (message has the cudaError) */
  PAID_EWORKSPACE = -4    /* workspace too small: see paid_attn_workspace_bytes */
} PaidStatus;

typedef enum PaidDType { PAID_F16 = 0, PAID_BF16 = 1 } PaidDType;

/* mode of one processor call */
typedef enum PaidMode {
  PAID_PLAIN = 0, /* deactivated processor: stock softmax attention of each frame        */
  PAID_OUTER = 1, /* (1-c) Attn(q,[k;k_begin]) + c Attn(q,[k;k_end])   interpolation.py:643-664 */
  PAID_INNER = 2  /* Attn(q,[k;(1-c)k_begin+c k_end])                   interpolation.py:772-790 */
} PaidMode;

/* flags */
#define PAID_FLAG_GENERIC_KERNELS 1u /* force the generic-shape CUDA kernels (validation cross-check) */
#define PAID_FLAG_ONE_WARPGROUP 2u   /* PLAIN / INNER: use the one-softmax-warpgroup tcgen05 kernel (attn_tc.cu) instead of
                                      * the persistent dual-warpgroup one (attn_dw.cu); cross-check and A/B timing */

typedef struct PaidAttnParams {
  uint32_t struct_size; /* sizeof(PaidAttnParams), ABI guard */
  uint32_t flags;
  int32_t dtype;  /* PaidDType of x, ctx, weights, y and the workspace tensors */
  int32_t mode;   /* PaidMode */
  int32_t fused;  /* 1: keys are [frame's own ; endpoint] ("fused with self-attention", is_fused) */
  int32_t N;      /* frames in this (local) batch */
  int32_t S;      /* query tokens per frame */
  int32_t L;      /* context tokens per frame (== S when ctx is NULL) */
  int32_t C;      /* channels (inner dim = heads * head_dim) */
  int32_t Cc;     /* context channels (== C when ctx is NULL) */
  int32_t heads;
  float scale;    /* attn.scale = head_dim^-0.5 */
  /* local indices of the two endpoint frames inside this batch (reference: 0 and N-1),
   * or -1 when that frame lives on another rank (then kv_ext must be given) */
  int32_t begin_frame;
  int32_t end_frame;
  const void* x;    /* (N,S,C)  hidden_states */
  const void* ctx;  /* NULL (self-attention) or (N,L,Cc) encoder_hidden_states */
  const void* wq;   /* (C,C)   attn.to_q.weight */
  const void* wk;   /* (C,Cc)  attn.to_k.weight */
  const void* wv;   /* (C,Cc)  attn.to_v.weight */
  const void* wo;   /* (C,C)   attn.to_out[0].weight */
  const void* bo;   /* (C,) or NULL  attn.to_out[0].bias */
  const float* coef; /* (N,) fp32 interpolation coefficients c_n of the local frames (ignored for PLAIN) */
  /* optional endpoint K/V produced elsewhere: (4,L,C) = K_begin, V_begin, K_end, V_end; NULL: take
   * them from frames begin_frame / end_frame of this batch */
  const void* kv_ext;
  void* y;          /* (N,S,C) output */
  void* workspace;  /* >= paid_attn_workspace_bytes(p) bytes, 256-byte aligned */
  uint64_t workspace_bytes;
  /* ---- ABI 2 ----
   * K and V of the context produced earlier (paid_attn_project_kv): the text prompt of a cross-attention layer does not
   * change over the denoising loop (pipeline_interpolated_sdxl.py:2232-2345 passes the same prompt_embeds every step),
   * so the caller projects it once per sequence.  (N,L,C) each, or one (L,C) matrix shared by every frame when
   * kv_pre_broadcast is 1 (PLAIN mode only: the unconditional pass, whose N frames carry the same negative prompt).
   * NULL: project ctx (or x) with wk / wv inside this call. */
  const void* k_pre;
  const void* v_pre;
  int32_t kv_pre_broadcast;
  /* ---- ABI 3 ----
   * Classifier-free-guidance rows: x, ctx (or k_pre / v_pre), y and the workspace carry N + plain_tail frames; frames
   * [0, N) are the interpolation sequence (mode / fused / coef / endpoints as above), frames [N, N + plain_tail) get stock
   * attention (PLAIN) in the same call -- the unconditional pass of the step (pipeline_interpolated_sdxl.py:2272-2293
   * runs it as a second UNet call with AID switched off).  The projections run once over all rows.  0: none. */
  int32_t plain_tail;
  /* cudaEvent_t (or NULL): the stream waits for it after the local projections and before the attention core -- the
   * endpoint K/V in kv_ext are being delivered on another stream (the NCCL broadcast of a frame-sharded sequence),
   * so the transfer overlaps the q/k/v projection of the local frames. */
  void* kv_ext_ready_event;
} PaidAttnParams;

/* attention on already projected tensors (head h of frame n lives at [n, t, h*d : (h+1)*d]) */
typedef struct PaidCoreParams {
  uint32_t struct_size;
  uint32_t flags;
  int32_t dtype, mode, fused;
  int32_t N, S, L, heads, head_dim;
  float scale;
  int32_t begin_frame, end_frame;
  const void* q;        /* (N,S,heads*head_dim) */
  const void* k;        /* (N,L,heads*head_dim) */
  const void* v;        /* (N,L,heads*head_dim) */
  const void* kv_ext;   /* NULL or (4,L,heads*head_dim) */
  const float* coef;    /* (N,) */
  void* out;            /* (N,S,heads*head_dim) attention output before to_out */
  void* workspace;      /* INNER only: 2*N*L*heads*head_dim elements for the lerped K/V; else may be NULL */
  uint64_t workspace_bytes;
  /* output combination (IP-Adapter variants, interpolation.py:364-367, 530, 196):
   *   out = (accumulate ? out : 0) + out_scale * (out_frame_scale ? out_frame_scale[n] : 1) * attention */
  int32_t accumulate;
  float out_scale;              /* 0 is read as 1 (zero-initialised structs keep the plain behaviour) */
  const float* out_frame_scale; /* NULL or (N,) fp32 */
  int32_t kv_broadcast;         /* 1: k and v are one (L, C) matrix shared by all frames (PLAIN mode only) */
} PaidCoreParams;

int paid_attn_abi_version(void);

/* bytes of workspace paid_attn_forward needs for these sizes (0 on invalid params) */
uint64_t paid_attn_workspace_bytes(const PaidAttnParams* p);
uint64_t paid_attn_core_workspace_bytes(const PaidCoreParams* p);

/* one processor call: q/k/v projection, interpolated attention, alpha-lerp, output projection */
int paid_attn_forward(const PaidAttnParams* p, void* cuda_stream);

/* the attention core alone */
int paid_attn_core(const PaidCoreParams* p, void* cuda_stream);

/* K and V of one local frame: k_out, v_out are (L,C).  Uses p->x / p->ctx, p->wk, p->wv. */
int paid_attn_project_endpoints(const PaidAttnParams* p, int32_t local_frame, void* k_out, void* v_out,
                                void* cuda_stream);

/* K and V of all N frames of the context (interpolation.py:623-624): k_out, v_out are (N,L,C), to be handed back through
 * PaidAttnParams.k_pre / v_pre on later calls.  Uses p->x / p->ctx, p->wk, p->wv, p->N, p->L, p->C, p->Cc. */
int paid_attn_project_kv(const PaidAttnParams* p, void* k_out, void* v_out, void* cuda_stream);

/* y (M,Nout) = x (M,K) * w(Nout,K)^T + bias(Nout or NULL) */
int paid_linear(const void* x, const void* w, const void* bias, void* y, int64_t M, int32_t Nout, int32_t K,
                int32_t dtype, uint32_t flags, void* cuda_stream);

/* The feed-forward's first Linear with its GEGLU in the GEMM epilogue (diffusers FeedForward / GEGLU: proj = Linear(C, 8C),
 * hidden, gate = proj(x).chunk(2); out = hidden * gelu(gate)):  w is (2 D, K) = [Wa ; Wg], bias (2 D,) = [ba ; bg] or NULL,
 *   y (M, D) = (x Wa^T + ba) * gelu(x Wg^T + bg)        (exact erf GELU, fp32 before the one rounding to the 16-bit output)
 * the (M, 2 D) intermediate never reaches HBM. */
int paid_linear_geglu(const void* x, const void* w, const void* bias, void* y, int64_t M, int32_t D, int32_t K,
                      int32_t dtype, uint32_t flags, void* cuda_stream);

/* GEGLU of the transformer block's feed-forward (the caller next to the attention path):
 * h is (M, 2*D) = [a | g] per row (output of the first FF Linear), out (M, D) = a * gelu(g), exact (erf) GELU. */
int paid_geglu(const void* h, void* out, int64_t M, int32_t D, int32_t dtype, void* cuda_stream);

/* Residual add + LayerNorm in front of every attention / feed-forward call of the transformer block (the caller of
 * the path; diffusers BasicTransformerBlock: x = x + attn(norm(x)) ...):
 *   x_out (rows,C) = x + delta          (skipped when delta is NULL; x_out may alias x or delta)
 *   h_out (rows,C) = LayerNorm(x_out) * gamma + beta     (statistics in fp32 over the C channels of a row)
 * C must be a multiple of 8 and at most 2048. */
int paid_add_layer_norm(const void* x, const void* delta, const void* gamma, const void* beta, void* x_out, void* h_out,
                        int64_t rows, int32_t C, float eps, int32_t dtype, void* cuda_stream);

/* GroupNorm (+ optional SiLU) of a channels-last feature map: x, y are (N, HW, C) = NHWC, gamma / beta (C,);
 * statistics per (frame, group) over HW x (C / groups) elements in fp32, fixed reduction order.
 * pre_bias: NULL or (N, C), added to x on load (the ResNet block's time-embedding add in front of its second norm).
 * workspace: >= paid_group_norm_workspace_bytes(...) bytes of caller-owned scratch for the partial statistics.
 * C % 8 == 0, C <= 4096, groups <= 64, C % groups == 0. */
uint64_t paid_group_norm_workspace_bytes(int32_t N, int64_t HW, int32_t C, int32_t groups);
int paid_group_norm_nhwc(const void* x, const void* pre_bias, const void* gamma, const void* beta, void* y, void* workspace,
                         uint64_t workspace_bytes, int32_t N, int64_t HW, int32_t C, int32_t groups, float eps,
                         int32_t silu, int32_t dtype, void* cuda_stream);

/* out (rows,C) = a + b + bias(C): the ResNet block's output -- shortcut + second conv (run without its bias) + the
 * conv biases -- in one pass instead of a broadcast bias add and a residual add.  out may alias a or b. */
int paid_residual_bias_add(const void* a, const void* b, const void* bias, void* out, int64_t rows, int32_t C,
                           int32_t dtype, void* cuda_stream);

/* message for the last non-OK status returned on this thread ("" if none) */
const char* paid_attn_last_error(void);

/* number of CUDA kernels this library has launched in this process (monotonic) */
uint64_t paid_attn_launch_count(void);

/* name of the attention kernel family the last paid_attn_core / paid_attn_forward call on this thread used:
 * "tcgen05", "tcgen05-padded" (head_dim < 64 zero-padded to 64 by the TMA unit) or "generic" ("" before the first call) */
const char* paid_attn_last_kernel(void);

/* Measurement hook (bench.py roofline): while enabled, every attention-core kernel launched by
 * paid_attn_forward / paid_attn_core is bracketed by CUDA events on the launching stream.
 * paid_attn_profile_read synchronises those events and returns, for the launches since the last reset, the
 * summed kernel time, the launch count and the summed ALGORITHMIC flops (SURVEY.md section 8d:
 * A = 2 N S L C per partial attention; outer-fused 6A, outer-pure 4A, inner-fused 4A, inner-pure / plain 2A). */
int paid_attn_profile_enable(int on);
int paid_attn_profile_read(double* total_ms, uint64_t* launches, double* alg_flops, int reset);

/* The same hook per kernel shape: while enabled, the projection / feed-forward GEMM launches are bracketed too, and
 * paid_attn_profile_rows returns one row per distinct (kind, shape) since the last reset:
 *   PAID_PROFILE_ATTENTION     d = {S, L, heads * head_dim, 16 * mode + fused}, flops = algorithmic flops (as above)
 *   PAID_PROFILE_LINEAR        d = {M, Nout, K, weight groups in the launch},  flops = 2 M Nout K groups
 *   PAID_PROFILE_LINEAR_GEGLU  d = {M, D, K, 1},                                flops = 4 M D K
 * At most `capacity` rows are written to `out`; *count receives the number of rows that exist (call with out == NULL
 * and capacity == 0 to size the array; that call does not reset). */
enum { PAID_PROFILE_ATTENTION = 0, PAID_PROFILE_LINEAR = 1, PAID_PROFILE_LINEAR_GEGLU = 2 };
typedef struct PaidProfileRow {
  int32_t kind;
  int32_t reserved;
  int64_t d[4];
  uint64_t launches;
  double total_ms;
  double flops;
} PaidProfileRow;
int paid_attn_profile_rows(PaidProfileRow* out, uint64_t capacity, uint64_t* count, int reset);

#ifdef __cplusplus
}
#endif
#endif /* PAID_ATTN_H_ */
