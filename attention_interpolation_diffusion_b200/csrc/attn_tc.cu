#include "paid_common.cuh"
namespace paid {
bool attn_tc_supported(const CoreArgs&) { return false; }
int launch_attn_tc(const CoreArgs&, cudaStream_t) { return fail(PAID_EUNSUPPORTED, "tcgen05 attention not built"); }
}  // namespace paid
