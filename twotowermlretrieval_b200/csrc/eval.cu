// eval.cu — rank of a designated document per query, without materialising [B, N].
//
// Replaces the metric loop of `BatchEvaluator.evaluate` (backend/evaluators.py:49-73): there the
// full similarity matrix is built with `torch.matmul`, every row is sorted with `torch.sort`, and the
// position of the positive document is searched.  The rank only needs a count:
//     rank_i = 1 + #{ j : s_ij > s_it  or  (s_ij == s_it and j < t) },   t = target[i]
// (the position in a stable descending sort).  One CTA per query streams the document rows once;
// every score, the target's included, is produced by the same fp32 summation tree, so comparisons of
// a score with itself are exact.
#include "common.cuh"

namespace ttr {

constexpr int PR_THREADS = 256;

__device__ __forceinline__ float row_dot(const float* __restrict__ d, const float* q_sm, int D, int lane) {
  float acc = 0.f;
  for (int c = lane * 4; c < D; c += 128) {
    const float4 v = *reinterpret_cast<const float4*>(d + c);
    acc = fmaf(v.x, q_sm[c], acc);
    acc = fmaf(v.y, q_sm[c + 1], acc);
    acc = fmaf(v.z, q_sm[c + 2], acc);
    acc = fmaf(v.w, q_sm[c + 3], acc);
  }
  return warp_sum(acc);
}

__global__ void __launch_bounds__(PR_THREADS)
positive_rank_kernel(const float* __restrict__ Q, const float* __restrict__ docs, const int64_t* __restrict__ target,
                     int B, int64_t N, int D, int32_t* __restrict__ rank_out, float* __restrict__ score_out) {
  extern __shared__ float q_sm[];
  __shared__ int counts[PR_THREADS / 32];
  const int q = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int c = threadIdx.x; c < D; c += PR_THREADS) q_sm[c] = Q[(int64_t)q * D + c];
  __syncthreads();
  const int64_t t = target[q];
  const bool has_t = t >= 0 && t < N;
  const float st = has_t ? row_dot(docs + t * D, q_sm, D, lane) : INFINITY;
  int cnt = 0;
  for (int64_t j = warp; j < N; j += PR_THREADS / 32) {
    const float s = row_dot(docs + j * D, q_sm, D, lane);
    cnt += (s > st || (s == st && j < t)) ? 1 : 0;
  }
  if (lane == 0) counts[warp] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    int tot = 0;
    for (int w = 0; w < PR_THREADS / 32; ++w) tot += counts[w];
    rank_out[q] = has_t ? tot + 1 : -1;
    if (score_out) score_out[q] = st;
  }
}

}  // namespace ttr

extern "C" int ttr_positive_rank(const float* Q, const float* docs, const int64_t* target, int B, int64_t N, int D,
                                 int32_t* rank_out, float* score_out, void* stream) {
  using namespace ttr;
  TTR_REQUIRE(B >= 1 && N >= 1 && D >= 4 && D % 4 == 0 && D <= 8192, "ttr_positive_rank: bad shape (B=%d N=%lld D=%d)", B,
              (long long)N, D);
  positive_rank_kernel<<<B, PR_THREADS, D * sizeof(float), (cudaStream_t)stream>>>(Q, docs, target, B, N, D, rank_out,
                                                                                 score_out);
  TTR_CHECK_LAUNCH();
  return TTR_OK;
}
