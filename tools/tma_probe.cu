// tma_probe.cu — micro-benchmark: how fast can one CTA per SM stream a row-major fp32 [N,256]
// matrix into shared memory with different TMA access shapes?  (Design input for
// score_topk_mma.cu; results are recorded in profiles/.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_probe tma_probe.cu -I../twotowermlretrieval_b200/csrc -I../include
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include "ptx.cuh"
using namespace ttr;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

constexpr int DIMF = 256;

// mode 0: 8 x 2D boxes {32 floats, rows}; mode 1: one 3D box {32, rows, 8}; mode 2: 1D bulk copy
template <int MODE>
__global__ void __launch_bounds__(64, 1)
probe(const __grid_constant__ CUtensorMap map2, const __grid_constant__ CUtensorMap map3, const float* docs,
      int64_t n_tiles, int rows, int stages, unsigned long long* sink) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* base = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  const int stage_bytes = rows * DIMF * 4;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(base + (size_t)stages * stage_bytes);
  uint64_t* empty_bar = full_bar + stages;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { ptx::mbar_init(full_bar + s, 1); ptx::mbar_init(empty_bar + s, 1); }
    ptx::fence_mbar_init();
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    int it = 0;
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
      const int s = it % stages; const uint32_t ph = (it / stages) & 1;
      ptx::mbar_wait(empty_bar + s, ph ^ 1u);
      ptx::mbar_arrive_expect_tx(full_bar + s, stage_bytes);
      unsigned char* dst = base + (size_t)s * stage_bytes;
      const int32_t d0 = (int32_t)(t * rows);
      if (MODE == 0) {
        for (int kb = 0; kb < 8; ++kb) ptx::tma_load_2d(dst + kb * rows * 128, &map2, kb * 32, d0, full_bar + s);
      } else if (MODE == 1) {
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::
                "r"(ptx::smem_u32(dst)), "l"(&map3), "r"(ptx::smem_u32(full_bar + s)), "r"(0), "r"(d0), "r"(0) : "memory");
      } else {
        ptx::bulk_g2s(dst, docs + (int64_t)d0 * DIMF, stage_bytes, full_bar + s);
      }
    }
  } else if (warp == 1 && lane == 0) {
    int it = 0; unsigned long long acc = 0;
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
      const int s = it % stages; const uint32_t ph = (it / stages) & 1;
      ptx::mbar_wait(full_bar + s, ph);
      acc += *reinterpret_cast<volatile unsigned int*>(base + (size_t)s * stage_bytes);
      ptx::mbar_arrive(empty_bar + s);
    }
    if (acc == 0x12345) *sink = acc;
  }
}

static PFN_cuTensorMapEncodeTiled_v12000 enc_fn() {
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
  return (PFN_cuTensorMapEncodeTiled_v12000)p;
}

template <int MODE>
float run(const CUtensorMap& m2, const CUtensorMap& m3, const float* docs, int64_t n, int rows, int stages, unsigned long long* sink) {
  size_t smem = (size_t)stages * rows * DIMF * 4 + 2 * stages * 8 + 1024;
  CK(cudaFuncSetAttribute(probe<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int64_t n_tiles = n / rows;
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  probe<MODE><<<148, 64, smem>>>(m2, m3, docs, n_tiles, rows, stages, sink);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int i = 0; i < 3; ++i) probe<MODE><<<148, 64, smem>>>(m2, m3, docs, n_tiles, rows, stages, sink);
  CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  return ms / 3;
}

int main() {
  const int64_t n = 4 * 1024 * 1024;   // 4 GiB of rows
  float* docs; CK(cudaMalloc(&docs, n * DIMF * 4)); CK(cudaMemset(docs, 0, n * DIMF * 4));
  unsigned long long* sink; CK(cudaMalloc(&sink, 8));
  auto enc = enc_fn();
  const int rows_list[] = {32, 64, 128};
  for (int rows : rows_list) {
    CUtensorMap m2, m3;
    {
      cuuint64_t dims[2] = {DIMF, (cuuint64_t)n}; cuuint64_t str[1] = {DIMF * 4};
      cuuint32_t box[2] = {32, (cuuint32_t)rows}; cuuint32_t es[2] = {1, 1};
      CUresult r = enc(&m2, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 2, docs, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r) { printf("enc2 failed %d\n", r); return 1; }
    }
    {
      cuuint64_t dims[3] = {32, (cuuint64_t)n, 8}; cuuint64_t str[2] = {DIMF * 4, 128};
      cuuint32_t box[3] = {32, (cuuint32_t)rows, 8}; cuuint32_t es[3] = {1, 1, 1};
      CUresult r = enc(&m3, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 3, docs, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r) { printf("enc3 failed %d (rows %d)\n", r, rows); }
    }
    const double gb = (double)n * DIMF * 4 / 1e9;
    for (int stages : {2, 3, 4, 6}) {
      if ((size_t)stages * rows * 1024 > 200 * 1024) continue;
      float a = run<0>(m2, m3, docs, n, rows, stages, sink);
      float b = run<1>(m2, m3, docs, n, rows, stages, sink);
      float c = run<2>(m2, m3, docs, n, rows, stages, sink);
      printf("rows %3d stages %d | 8x2D boxes %7.1f GB/s | 1x3D box %7.1f GB/s | 1D bulk %7.1f GB/s\n", rows, stages,
             gb / (a * 1e-3), gb / (b * 1e-3), gb / (c * 1e-3));
    }
  }
  return 0;
}
