// Host-side TMA descriptor construction.  cuTensorMapEncodeTiled is fetched through the runtime's
// driver-entry-point API so the library has no link-time dependency on libcuda (it must load on a
// GPU-less build box).  Encoded descriptors are cached by (pointer, geometry).
#include <mutex>
#include <unordered_map>

#include "paid_common.cuh"
#include "sm100_ptx.cuh"

namespace paid {

using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                              const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                              CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeFn encode_fn() {
  static EncodeFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return (EncodeFn)p;
  }();
  return fn;
}

struct Key {
  const void* base;
  long long a, b, c, d, e;
  int dtype, box, kind;
  bool operator==(const Key& o) const {
    return base == o.base && a == o.a && b == o.b && c == o.c && d == o.d && e == o.e && dtype == o.dtype &&
           box == o.box && kind == o.kind;
  }
};
struct KeyHash {
  size_t operator()(const Key& k) const {
    size_t h = std::hash<const void*>()(k.base);
    auto mix = [&h](long long v) { h ^= std::hash<long long>()(v) + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2); };
    mix(k.a); mix(k.b); mix(k.c); mix(k.d); mix(k.e); mix(k.dtype); mix(k.box); mix(k.kind);
    return h;
  }
};
static std::mutex g_mu;
static std::unordered_map<Key, CUtensorMap, KeyHash> g_cache;

static int encode(CUtensorMap* out, const Key& key, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
                  const cuuint32_t* box) {
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_cache.find(key);
    if (it != g_cache.end()) { *out = it->second; return PAID_OK; }
  }
  EncodeFn fn = encode_fn();
  if (!fn) return fail(PAID_ECUDA, "cuTensorMapEncodeTiled is unavailable (no CUDA driver?)");
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(out, key.dtype == PAID_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                  (cuuint32_t)rank, const_cast<void*>(key.base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(PAID_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_cache.size() > 8192) g_cache.clear();
  g_cache.emplace(key, *out);
  return PAID_OK;
}

int make_tmap_2d(CUtensorMap* out, const void* base, int dtype, long long rows, long long cols, long long row_pitch,
                 int box_rows) {
  Key key{base, rows, cols, row_pitch, 0, 0, dtype, box_rows, 2};
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)row_pitch * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  return encode(out, key, 2, dims, strides, box);
}

int make_tmap_heads(CUtensorMap* out, const void* base, int dtype, long long frames, long long tokens, int heads,
                    int head_dim, long long frame_stride_elems, int box_tokens) {
  Key key{base, frames, tokens, heads, head_dim, frame_stride_elems, dtype, box_tokens, 4};
  const long long C = (long long)heads * head_dim;
  cuuint64_t dims[4] = {(cuuint64_t)head_dim, (cuuint64_t)heads, (cuuint64_t)tokens, (cuuint64_t)frames};
  // a single shared (L, C) matrix is described as one frame
  cuuint64_t strides[3] = {(cuuint64_t)head_dim * 2, (cuuint64_t)C * 2,
                           (cuuint64_t)(frame_stride_elems > 0 ? frame_stride_elems : tokens * C) * 2};
  cuuint32_t box[4] = {64, 1, (cuuint32_t)box_tokens, 1};
  return encode(out, key, 4, dims, strides, box);
}

}  // namespace paid
