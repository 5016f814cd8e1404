// mma_probe.cu — micro-benchmark: cycles per tcgen05.mma kind::tf32 (M = 128, K = 8) as a
// function of N, operand source (A from TMEM or shared memory) and the number of independent
// accumulation chains.  Design input for score_topk_mma.cu.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include "ptx.cuh"
using namespace ttr;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

template <int N, int CHAINS, bool A_TMEM>
__global__ void __launch_bounds__(128, 1) probe(long long* out, int rounds) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* base = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<float*>(base)[i] = 0.001f * (i & 63);
  if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::fence_mbar_init(); }
  if (threadIdx.x < 32) { ptx::tmem_alloc(&slot, 512); ptx::tmem_relinquish(); }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tm = slot;
  if (threadIdx.x < 32) {
    constexpr uint32_t idesc = ptx::make_idesc_tf32(128, N);
    const uint64_t a_desc = ptx::make_kmajor_sw128_desc(ptx::smem_u32(base));
    const uint64_t b_desc = ptx::make_kmajor_sw128_desc(ptx::smem_u32(base) + 16384);
    long long t0 = clock64();
    for (int r = 0; r < rounds; ++r) {
      if (ptx::elect_one()) {
#pragma unroll
        for (int kk = 0; kk < 32 / CHAINS; ++kk) {
#pragma unroll
          for (int c = 0; c < CHAINS; ++c) {
            const int ks = c * (32 / CHAINS) + kk;
            if (A_TMEM) ptx::mma_tf32_ts(tm + 256 + c * (N <= 64 ? N : 0), tm + (ks & 31) * 8, b_desc + 2 * (ks & 3), idesc, kk != 0);
            else ptx::mma_tf32_ss(tm + 256 + c * (N <= 64 ? N : 0), a_desc + 2 * (ks & 3), b_desc + 2 * (ks & 3), idesc, kk != 0);
          }
        }
        ptx::mma_commit(&bar);
      }
      __syncwarp();
      ptx::mbar_wait(&bar, r & 1);
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) *out = t1 - t0;
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) ptx::tmem_dealloc(tm, 512);
}

template <int N, int CHAINS, bool A_TMEM>
void run(const char* name, long long* d_out) {
  const int rounds = 2000;
  size_t smem = 64 * 1024;
  CK(cudaFuncSetAttribute(probe<N, CHAINS, A_TMEM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  probe<N, CHAINS, A_TMEM><<<148, 128, smem>>>(d_out, rounds);
  CK(cudaDeviceSynchronize());
  long long cyc; CK(cudaMemcpy(&cyc, d_out, 8, cudaMemcpyDeviceToHost));
  printf("%-34s N=%3d chains=%d : %7.1f cycles per 32-MMA tile, %5.1f per MMA (incl. commit+wait per tile)\n", name, N, CHAINS,
         (double)cyc / rounds, (double)cyc / rounds / 32);
}

int main() {
  long long* d; CK(cudaMalloc(&d, 8));
  run<32, 1, true>("A=TMEM", d);
  run<32, 4, true>("A=TMEM", d);
  run<32, 1, false>("A=SMEM", d);
  run<32, 4, false>("A=SMEM", d);
  run<64, 1, true>("A=TMEM", d);
  run<64, 4, true>("A=TMEM", d);
  run<64, 1, false>("A=SMEM", d);
  run<128, 1, true>("A=TMEM", d);
  run<128, 1, false>("A=SMEM", d);
  run<256, 1, false>("A=SMEM", d);
  run<256, 1, true>("A=TMEM", d);
  return 0;
}
