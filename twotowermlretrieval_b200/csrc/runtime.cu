// runtime.cu — thread-local error text and device queries.
#include <stdarg.h>
#include <string.h>

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

// HOST helper of the input pipeline (no device work): ragged token rows -> one right-padded [n_rows, T] int64 matrix
// (`pad_sequence(batch_first=True, padding_value=0)`, backend/main.py:50-56) written straight into a pinned staging
// buffer.  Row r of the output is row rows[r] of the ragged set (flat ids + start offsets + lengths).
extern "C" int ttr_pack_padded_i64(const int64_t* flat, const int64_t* starts, const int64_t* lengths,
                                   const int64_t* rows, int64_t n_rows, int64_t T, int64_t* out) {
  TTR_REQUIRE(flat && starts && lengths && rows && out && n_rows >= 0 && T >= 0, "ttr_pack_padded_i64: bad arguments");
  for (int64_t r = 0; r < n_rows; ++r) {
    const int64_t src = rows[r];
    int64_t len = lengths[src];
    if (len > T) len = T;
    int64_t* dst = out + r * T;
    memcpy(dst, flat + starts[src], (size_t)len * sizeof(int64_t));
    memset(dst + len, 0, (size_t)(T - len) * sizeof(int64_t));
  }
  return TTR_OK;
}

// The same copy, counting on the way what the device plan will count again: `nnz_total` = non-zero ids in the packed
// rows (the number of packed tokens: SURVEY quirk #1, the effective length of a row is its count of non-zero ids),
// `zero_rows` = rows without any (the reference's pack_padded_sequence raises for those, quirk #2).  One pass over the
// tokens instead of a separate numpy pass over the whole corpus before the first batch can start (0.9 s per 1.1 M
// passages on the host that measured it).
extern "C" int ttr_pack_padded_count_i64(const int64_t* flat, const int64_t* starts, const int64_t* lengths,
                                         const int64_t* rows, int64_t n_rows, int64_t T, int64_t* out,
                                         int64_t* nnz_total, int64_t* zero_rows) {
  TTR_REQUIRE(flat && starts && lengths && rows && out && nnz_total && zero_rows && n_rows >= 0 && T >= 0,
              "ttr_pack_padded_count_i64: bad arguments");
  int64_t total = 0, empty = 0;
  for (int64_t r = 0; r < n_rows; ++r) {
    const int64_t src = rows[r];
    int64_t len = lengths[src];
    if (len > T) len = T;
    if (len < 0) len = 0;
    int64_t* dst = out + r * T;
    const int64_t* from = flat + starts[src];
    int64_t nz = 0;
    for (int64_t j = 0; j < len; ++j) {
      const int64_t v = from[j];
      dst[j] = v;
      nz += v != 0;
    }
    memset(dst + len, 0, (size_t)(T - len) * sizeof(int64_t));
    total += nz;
    empty += nz == 0;
  }
  *nnz_total = total;
  *zero_rows = empty;
  return TTR_OK;
}
