// dropout.cu — inter-layer dropout of nn.GRU (backend/model.py:35): scaled keep mask on the
// per-step outputs of every layer but the last, applied only in train mode.  Counter-based
// generator (one 64-bit mix per element of (seed, index)), so the mask is reproducible from the
// seed and can be exported for replay through the oracle.
#include "common.cuh"

namespace ttr {

__device__ __forceinline__ uint32_t mix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  x = x ^ (x >> 31);
  return (uint32_t)(x >> 32);
}

__global__ void dropout_kernel(const float* __restrict__ y, int64_t n, float p, float scale, uint64_t seed,
                               float* __restrict__ out, float* __restrict__ mask) {
  const uint32_t thresh = (uint32_t)((double)p * 4294967296.0);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float m = mix64(seed * 0xD1342543DE82EF95ull + (uint64_t)i) >= thresh ? scale : 0.f;
    mask[i] = m;
    out[i] = y[i] * m;
  }
}

}  // namespace ttr

extern "C" int ttr_dropout(const float* y, int64_t n, float p, uint64_t seed, float* out, float* mask, void* stream) {
  using namespace ttr;
  TTR_REQUIRE(n >= 1 && p >= 0.f && p < 1.f, "ttr_dropout: bad arguments (n=%lld, p=%f)", (long long)n, p);
  const int grid = (int)std::min<int64_t>(8 * sm_count(), ceil_div64(n, 256));
  dropout_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(y, n, p, 1.0f / (1.0f - p), seed, out, mask);
  TTR_CHECK_LAUNCH();
  return TTR_OK;
}
