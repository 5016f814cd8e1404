// gemm_fp32.cu — plain fp32 CUDA-core GEMMs with arbitrary operand strides.
//
// Used for (a) the gradient contractions of the GRU backward (dX = dG W, dW = dG^T X; the
// autograd of backend/main.py:254) where fp32 accumulation order, not tensor-core speed, is
// what the parity tests pin, and (b) `ttr_debug_gemm_fp32_bias`, the test-only reference of
// the tcgen05 input-projection kernel.  64x64x16 tiles, 4x4 register blocking.
#include "common.cuh"

namespace ttr {

constexpr int TM = 64, TN = 64, TK = 16;

struct GemmArgs {
  const float* A; int64_t sa_i, sa_l;
  const float* B; int64_t sb_l, sb_j;
  const float* bias;
  float* C; int64_t ldc;
  int I, J, L;
  const int32_t* dyn; int dyn_which;   // 1: I = *dyn, 2: L = *dyn
  int accumulate;
};

__global__ void __launch_bounds__(256) gemm_fp32_kernel(GemmArgs g) {
  __shared__ float As[TK][TM + 4];
  __shared__ float Bs[TK][TN + 4];
  int I = g.I, L = g.L;
  if (g.dyn) {
    int v = *g.dyn;
    if (g.dyn_which == 1) I = min(I, v);
    if (g.dyn_which == 2) L = min(L, v);
  }
  const int i0 = blockIdx.y * TM, j0 = blockIdx.x * TN;
  if (i0 >= I) return;
  // split of the reduction dimension across gridDim.z
  const int l_per = ceil_div(ceil_div(L, (int)gridDim.z), TK) * TK;
  const int l_begin = blockIdx.z * l_per;
  const int l_end = min(L, l_begin + l_per);
  if (l_begin >= l_end && !(blockIdx.z == 0)) return;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const bool a_l_contig = (g.sa_l == 1);
  const bool b_l_contig = (g.sb_l == 1);
  float acc[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;

  for (int l0 = l_begin; l0 < l_end; l0 += TK) {
#pragma unroll
    for (int e4 = 0; e4 < 4; ++e4) {
      const int e = threadIdx.x + e4 * 256;
      int i, l;
      if (a_l_contig) { i = e >> 4; l = e & 15; } else { i = e & 63; l = e >> 6; }
      float v = 0.f;
      if (i0 + i < I && l0 + l < l_end) v = g.A[(int64_t)(i0 + i) * g.sa_i + (int64_t)(l0 + l) * g.sa_l];
      As[l][i] = v;
      int j, lb;
      if (b_l_contig) { j = e >> 4; lb = e & 15; } else { j = e & 63; lb = e >> 6; }
      float wv = 0.f;
      if (j0 + j < g.J && l0 + lb < l_end) wv = g.B[(int64_t)(l0 + lb) * g.sb_l + (int64_t)(j0 + j) * g.sb_j];
      Bs[lb][j] = wv;
    }
    __syncthreads();
#pragma unroll
    for (int l = 0; l < TK; ++l) {
      const float4 av = *reinterpret_cast<const float4*>(&As[l][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[l][tx * 4]);
      const float ar[4] = {av.x, av.y, av.z, av.w};
      const float br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = fmaf(ar[r], br[c], acc[r][c]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int i = i0 + ty * 4 + r;
    if (i >= I) continue;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int j = j0 + tx * 4 + c;
      if (j >= g.J) continue;
      float v = acc[r][c];
      if (g.bias && blockIdx.z == 0) v += g.bias[j];
      float* dst = g.C + (int64_t)i * g.ldc + j;
      if (gridDim.z > 1) atomicAdd(dst, v);
      else if (g.accumulate) *dst += v;
      else *dst = v;
    }
  }
}

static int launch_gemm(GemmArgs g, int split, cudaStream_t st) {
  if (split > 1 && !g.accumulate) {
    TTR_REQUIRE(g.ldc == g.J, "gemm split needs a dense C");
    TTR_CHECK_CUDA(cudaMemsetAsync(g.C, 0, (size_t)g.I * g.J * sizeof(float), st));
  }
  dim3 grid(ceil_div(g.J, TN), ceil_div(g.I, TM), split);
  gemm_fp32_kernel<<<grid, 256, 0, st>>>(g);
  TTR_CHECK_LAUNCH();
  return TTR_OK;
}

int launch_gemm_strided(const float* A, int64_t sa_i, int64_t sa_l, const float* Bm, int64_t sb_l, int64_t sb_j,
                        float* C, int64_t ldc, int I, int J, int L, const int32_t* dyn, int dyn_which,
                        int accumulate, int allow_split, cudaStream_t st) {
  GemmArgs g{A, sa_i, sa_l, Bm, sb_l, sb_j, nullptr, C, ldc, I, J, L, dyn, dyn_which, accumulate};
  int split = 1;
  if (allow_split && L > 4096) {
    int tiles = ceil_div(I, TM) * ceil_div(J, TN);
    split = max(1, min(64, (2 * sm_count()) / max(tiles, 1)));
    split = min(split, ceil_div(L, 1024));
  }
  return launch_gemm(g, split, st);
}

}  // namespace ttr

extern "C" int ttr_debug_gemm_fp32_bias(const float* A, const float* W, const float* bias, float* C,
                                        int m_bound, const int32_t* m_valid, int N, int K, void* stream) {
  using namespace ttr;
  TTR_REQUIRE(m_bound >= 1 && N >= 1 && K >= 1, "ttr_debug_gemm_fp32_bias: bad shape");
  GemmArgs g{A, K, 1, W, 1, K, bias, C, N, m_bound, N, K, m_valid, 1, 0};
  return launch_gemm(g, 1, (cudaStream_t)stream);
}

extern "C" int ttr_gemm_nn_fp32(const float* A, const float* W, float* C, int m_bound, const int32_t* m_valid,
                                int N, int K, int accumulate, void* stream) {
  using namespace ttr;
  TTR_REQUIRE(m_bound >= 1 && N >= 1 && K >= 1, "ttr_gemm_nn_fp32: bad shape");
  // C[M,K] = A[M,N] * W[N,K]: reduction over N
  GemmArgs g{A, N, 1, W, K, 1, nullptr, C, K, m_bound, K, N, m_valid, 1, accumulate};
  return launch_gemm(g, 1, (cudaStream_t)stream);
}

extern "C" int ttr_gemm_tn_fp32(const float* A, const float* Bm, float* C, int m_bound, const int32_t* m_valid,
                                int N, int K, int accumulate, void* stream) {
  using namespace ttr;
  TTR_REQUIRE(m_bound >= 1 && N >= 1 && K >= 1, "ttr_gemm_tn_fp32: bad shape");
  // C[N,K] = A[M,N]^T * Bm[M,K]: reduction over M (dynamic), split across CTAs
  GemmArgs g{A, 1, N, Bm, K, 1, nullptr, C, K, N, K, m_bound, m_valid, 2, accumulate};
  int tiles = ceil_div(N, TM) * ceil_div(K, TN);
  int split = 1;
  if (m_bound > 4096) {
    split = max(1, min(64, (2 * sm_count()) / max(tiles, 1)));
    split = min(split, ceil_div(m_bound, 1024));
  }
  return launch_gemm(g, split, (cudaStream_t)stream);
}
