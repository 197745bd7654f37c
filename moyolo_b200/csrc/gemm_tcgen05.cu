// placeholder until the tcgen05 engine lands (next commit)
#include "common.cuh"
namespace moyolo {
bool linear_tcgen05_supported(const void*, int64_t, const void*, int64_t, int, int) { return false; }
int linear_tcgen05(const void*, int64_t, const void*, const float*, void*, int64_t, int64_t, int, int, int, int,
                   const uint8_t*, cudaStream_t) {
  return fail(MOYOLO_ERR_UNSUPPORTED, "tcgen05 engine not built");
}
}  // namespace moyolo
