// hybrid.cu — TF-IDF hybrid reweighting of the dense top-k on the device.
//
// Replaces the Python loop of frontend/main.py:158-198: semantic = 1 - dist (Chroma's
// default squared-L2 space => 2*cos - 1; quirk #9), tfidf = cosine of L2-normalised TF-IDF
// rows = sparse dot, final = alpha*semantic + (1-alpha)*tfidf, stable descending sort,
// first top_n.  All blend arithmetic is fp64 with explicit round-to-nearest mul/add (no FMA
// contraction) so it reproduces the reference's Python-float evaluation order.
#include "common.cuh"

namespace ttr {

__device__ __forceinline__ double sparse_row_dot(const int64_t* __restrict__ indptr,
                                                 const int32_t* __restrict__ indices,
                                                 const double* __restrict__ data, int64_t row,
                                                 const int32_t* __restrict__ q_idx,
                                                 const double* __restrict__ q_val, int q_nnz) {
  // ascending feature order == scipy's CSR product accumulation order (sorted indices)
  double acc = 0.0;
  const int64_t lo = indptr[row], hi = indptr[row + 1];
  int qp = 0;
  for (int64_t e = lo; e < hi && qp < q_nnz; ++e) {
    const int32_t c = indices[e];
    while (qp < q_nnz && q_idx[qp] < c) ++qp;
    if (qp < q_nnz && q_idx[qp] == c) acc = __dadd_rn(acc, __dmul_rn(data[e], q_val[qp]));
  }
  return acc;
}

__global__ void tfidf_candidates_kernel(const int64_t* __restrict__ cand_idx, int B, int kc,
                                        int64_t row_offset, int64_t rows,
                                        const int64_t* __restrict__ indptr,
                                        const int32_t* __restrict__ indices,
                                        const double* __restrict__ data,
                                        const int64_t* __restrict__ q_indptr,
                                        const int32_t* __restrict__ q_indices,
                                        const double* __restrict__ q_data, double* __restrict__ out) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)B * kc) return;
  const int q = (int)(t / kc);
  const int64_t local = cand_idx[t] - row_offset;
  double v = 0.0;
  if (cand_idx[t] >= 0 && local >= 0 && local < rows) {
    const int64_t qlo = q_indptr[q];
    const int qn = (int)(q_indptr[q + 1] - qlo);
    if (qn > 0) v = sparse_row_dot(indptr, indices, data, local, q_indices + qlo, q_data + qlo, qn);
  }
  out[t] = v;
}

// one CTA per query, one thread per candidate (kc <= 128)
__global__ void __launch_bounds__(128)
hybrid_rerank_kernel(const int64_t* __restrict__ cand_idx, const float* __restrict__ cand_cos, int kc,
                     int64_t row_offset, const int64_t* __restrict__ indptr,
                     const int32_t* __restrict__ indices, const double* __restrict__ data,
                     const int64_t* __restrict__ q_indptr, const int32_t* __restrict__ q_indices,
                     const double* __restrict__ q_data, const double* __restrict__ tfidf_in,
                     const double* __restrict__ q_sqnorm, const double* __restrict__ d_sqnorm, double alpha,
                     int space, int top_n, double* __restrict__ out_final, double* __restrict__ out_sem,
                     double* __restrict__ out_tfidf, int32_t* __restrict__ out_pos) {
  __shared__ double fin[128];
  const int q = blockIdx.x, j = threadIdx.x;
  double sem = 0.0, tf = 0.0, f = -INFINITY;
  const bool valid = j < kc && cand_idx[(int64_t)q * kc + j] >= 0;
  if (valid) {
    const double c = (double)cand_cos[(int64_t)q * kc + j];
    if (space != 0) {
      sem = c;
    } else if (q_sqnorm == nullptr && d_sqnorm == nullptr) {
      sem = __dadd_rn(__dmul_rn(2.0, c), -1.0);                    // unit vectors: 1 - |q - d|^2 = 2 cos - 1
    } else {
      // general squared-L2 distance of Chroma's default space: 1 - (|q|^2 + |d|^2 - 2 q.d); a token-less query is the
      // zero vector (query_inferencer.py:65-69) -> dist = |d|^2, an un-normalised model has |d| != 1
      const double qn = q_sqnorm ? q_sqnorm[q] : 1.0;
      const double dn = d_sqnorm ? d_sqnorm[(int64_t)q * kc + j] : 1.0;
      sem = __dadd_rn(1.0, -__dadd_rn(__dadd_rn(qn, dn), -__dmul_rn(2.0, c)));
    }
    if (tfidf_in) {
      tf = tfidf_in[(int64_t)q * kc + j];
    } else {
      const int64_t qlo = q_indptr[q];
      const int qn = (int)(q_indptr[q + 1] - qlo);
      if (qn > 0)
        tf = sparse_row_dot(indptr, indices, data, cand_idx[(int64_t)q * kc + j] - row_offset,
                            q_indices + qlo, q_data + qlo, qn);
    }
    if (tf != tf) tf = 0.0;   // np.nan_to_num (frontend/main.py:172)
    f = __dadd_rn(__dmul_rn(alpha, sem), __dmul_rn(__dadd_rn(1.0, -alpha), tf));
  }
  fin[j] = f;
  __syncthreads();
  if (j < kc) {
    // stable descending rank == Python's list.sort(key=score, reverse=True)
    int rank = 0;
    for (int i = 0; i < kc; ++i) {
      const double o = fin[i];
      rank += (o > f) || (o == f && i < j);
    }
    if (rank < top_n) {
      const int64_t o = (int64_t)q * top_n + rank;
      out_final[o] = f;
      out_sem[o] = sem;
      out_tfidf[o] = tf;
      out_pos[o] = valid ? j : -1;
    }
  }
}

}  // namespace ttr

extern "C" int ttr_tfidf_candidates(const int64_t* cand_idx, int B, int kc, int64_t csr_row_offset,
                                    int64_t csr_rows, const int64_t* indptr, const int32_t* indices,
                                    const double* data, const int64_t* q_indptr, const int32_t* q_indices,
                                    const double* q_data, double* out, void* stream) {
  using namespace ttr;
  TTR_REQUIRE(B >= 1 && kc >= 1, "ttr_tfidf_candidates: empty problem");
  const int64_t n = (int64_t)B * kc;
  tfidf_candidates_kernel<<<(unsigned)ceil_div64(n, 128), 128, 0, (cudaStream_t)stream>>>(
      cand_idx, B, kc, csr_row_offset, csr_rows, indptr, indices, data, q_indptr, q_indices, q_data, out);
  TTR_CHECK_LAUNCH();
  return TTR_OK;
}

extern "C" int ttr_hybrid_rerank(const int64_t* cand_idx, const float* cand_cos, int B, int kc,
                                 int64_t csr_row_offset, const int64_t* indptr, const int32_t* indices,
                                 const double* data, const int64_t* q_indptr, const int32_t* q_indices,
                                 const double* q_data, const double* tfidf_in, const double* q_sqnorm,
                                 const double* d_sqnorm, double alpha, int space,
                                 int top_n, double* out_final, double* out_sem, double* out_tfidf,
                                 int32_t* out_pos, void* stream) {
  using namespace ttr;
  TTR_REQUIRE(B >= 1 && kc >= 1 && kc <= 128, "ttr_hybrid_rerank: kc=%d outside [1,128]", kc);
  TTR_REQUIRE(top_n >= 1 && top_n <= kc, "ttr_hybrid_rerank: top_n=%d outside [1,kc]", top_n);
  TTR_REQUIRE(tfidf_in != nullptr || (indptr && indices && data && q_indptr && q_indices && q_data),
              "ttr_hybrid_rerank: need either tfidf_in or the CSR operands");
  hybrid_rerank_kernel<<<B, 128, 0, (cudaStream_t)stream>>>(cand_idx, cand_cos, kc, csr_row_offset, indptr,
                                                           indices, data, q_indptr, q_indices, q_data,
                                                           tfidf_in, q_sqnorm, d_sqnorm, alpha, space, top_n, out_final, out_sem,
                                                           out_tfidf, out_pos);
  TTR_CHECK_LAUNCH();
  return TTR_OK;
}
