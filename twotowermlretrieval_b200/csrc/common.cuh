// common.cuh — error plumbing and small device helpers shared by all kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "ttr_b200.h"

namespace ttr {

void set_error(const char* fmt, ...);

#define TTR_CHECK_CUDA(expr)                                                          \
  do {                                                                                \
    cudaError_t _e = (expr);                                                          \
    if (_e != cudaSuccess) {                                                          \
      ::ttr::set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,           \
                       cudaGetErrorString(_e));                                       \
      return TTR_ERR_CUDA;                                                            \
    }                                                                                 \
  } while (0)

#define TTR_REQUIRE(cond, ...)                                                        \
  do {                                                                                \
    if (!(cond)) {                                                                    \
      ::ttr::set_error(__VA_ARGS__);                                                  \
      return TTR_ERR_INVALID;                                                         \
    }                                                                                 \
  } while (0)

#define TTR_CHECK_LAUNCH() TTR_CHECK_CUDA(cudaGetLastError())

int sm_count();

__host__ __device__ inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// streaming 128-bit load that does not allocate in L1 (data touched once)
__device__ __forceinline__ float4 ld_stream_f4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

__device__ __forceinline__ float round_tf32(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

// Total order used by every top-k stage: higher score first, then lower index.
__device__ __forceinline__ bool cand_better(float sa, int64_t ia, float sb, int64_t ib) {
  return sa > sb || (sa == sb && ia < ib);
}

}  // namespace ttr
