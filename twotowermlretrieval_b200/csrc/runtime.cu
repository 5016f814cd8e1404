// runtime.cu — thread-local error text and device queries.
#include <stdarg.h>

#include "common.cuh"

namespace ttr {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return 148;
    cached = prop.multiProcessorCount;
    cached_dev = dev;
  }
  return cached;
}

}  // namespace ttr

extern "C" const char* ttr_last_error(void) { return ttr::g_err; }
extern "C" int ttr_version(void) { return 100; }
extern "C" int ttr_sm_count(int* out_h) {
  int dev = 0;
  TTR_CHECK_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  TTR_CHECK_CUDA(cudaGetDeviceProperties(&prop, dev));
  TTR_REQUIRE(prop.major == 10, "ttr_b200 needs an sm_100 device, found sm_%d%d", prop.major, prop.minor);
  *out_h = prop.multiProcessorCount;
  return TTR_OK;
}
