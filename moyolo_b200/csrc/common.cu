// Status / error plumbing and small host helpers shared by every entry point.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace moyolo {

static thread_local char g_last_error[512] = "";
static thread_local unsigned long long g_launches = 0;

void note_launch() { ++g_launches; }
unsigned long long launches() { return g_launches; }

char* last_error_buf() { return g_last_error; }

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
  return code;
}

bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("MOYOLO_PDL");
    return !(e != nullptr && e[0] == '0');
  }();
  return on;
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(MOYOLO_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  return MOYOLO_OK;
}

int make_levels(const int32_t* shapes_hw_host, int n_levels, int64_t len_v, LevelTable* out) {
  MOYOLO_REQUIRE(n_levels >= 1 && n_levels <= MOYOLO_MAX_LEVELS, MOYOLO_ERR_BAD_ARG,
                 "n_levels must be in [1, %d], got %d", MOYOLO_MAX_LEVELS, n_levels);
  int64_t acc = 0;
  out->n = n_levels;
  for (int l = 0; l < n_levels; ++l) {
    const int h = shapes_hw_host[2 * l], w = shapes_hw_host[2 * l + 1];
    MOYOLO_REQUIRE(h > 0 && w > 0, MOYOLO_ERR_BAD_SHAPE, "level %d has non-positive shape (%d, %d)", l, h, w);
    out->h[l] = h;
    out->w[l] = w;
    out->start[l] = static_cast<int>(acc);
    acc += static_cast<int64_t>(h) * w;
  }
  MOYOLO_REQUIRE(acc == len_v, MOYOLO_ERR_BAD_SHAPE,
                 "sum of value_shapes H*W (%lld) does not match value length (%lld)", (long long)acc,
                 (long long)len_v);
  MOYOLO_REQUIRE(acc < (1ll << 31), MOYOLO_ERR_BAD_SHAPE, "value length exceeds int32 index range");
  return MOYOLO_OK;
}

}  // namespace moyolo

namespace moyolo { unsigned long long launches(); }
extern "C" int moyolo_version(void) { return MOYOLO_VERSION; }
extern "C" uint64_t moyolo_launch_count(void) { return moyolo::launches(); }
extern "C" const char* moyolo_last_error(void) { return moyolo::last_error_buf(); }
extern "C" int moyolo_device_supported(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return 0; }
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return major == 10 ? 1 : 0;
}
