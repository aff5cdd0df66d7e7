#include "paid_common.cuh"
namespace paid {
bool linear_tc_supported(long long, int, int) { return false; }
int launch_linear_tc(const void*, const void*, const void*, void*, long long, int, int, int, cudaStream_t) {
  return fail(PAID_EUNSUPPORTED, "tcgen05 linear not built");
}
}  // namespace paid
