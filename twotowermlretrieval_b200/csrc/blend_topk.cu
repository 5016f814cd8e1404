// blend_topk.cu — corpus-wide hybrid search: combined[i] = alpha * cos(q, d_i) + (1 - alpha) *
// tfidf_cos(q, d_i) over ALL documents, then the top_k best in `np.argsort(combined)[::-1]` order.
//
// Replaces `SimpleHybridRetriever.search` (backend/simple_hybrid.py:45-66): sklearn
// `cosine_similarity` on the dense embeddings (float32: normalise both sides, dot), the TF-IDF
// cosine against the whole CSR matrix (float64 sparse dot of L2-normalised rows), the blend
// `self.alpha * dense_scores + (1 - self.alpha) * tfidf_scores` (float32 product promoted to
// float64 by the sum) and the full argsort.  One streaming pass: a warp per document computes
// the dense dot with coalesced 128-bit loads and lane 0 walks the sparse row; every warp keeps
// a 128-slot (double score, id) buffer with bitonic compaction; a second kernel merges the
// per-CTA lists.  Ties: argsort()[::-1] puts the HIGHER index first, and so does this.
#include "common.cuh"

namespace ttr {

constexpr int BL_WARPS = 8;
constexpr int BL_CAP = 128;
constexpr int BL_KMAX = 64;

// total order of np.argsort(x)[::-1] on distinct ids: higher score first, then higher id
__device__ __forceinline__ bool blend_better(double sa, int ia, double sb, int ib) {
  return sa > sb || (sa == sb && ia > ib);
}

__device__ __forceinline__ void warp_bitonic_blend(double* s, int* ix, int n_slots, int lane) {
  for (int k = 2; k <= n_slots; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = lane; t < (n_slots >> 1); t += 32) {
        int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        int p = i | j;
        bool desc = ((i & k) == 0);
        double si = s[i], sp = s[p];
        int ii = ix[i], ip = ix[p];
        if (blend_better(sp, ip, si, ii) == desc) { s[i] = sp; s[p] = si; ix[i] = ip; ix[p] = ii; }
      }
      __syncwarp();
    }
  }
}

__device__ __forceinline__ int warp_compact_blend(double* s, int* ix, int cnt, int k, int lane, double& tau_s, int& tau_i) {
  for (int t = cnt + lane; t < BL_CAP; t += 32) { s[t] = -INFINITY; ix[t] = -1; }
  __syncwarp();
  warp_bitonic_blend(s, ix, BL_CAP, lane);
  const int kept = cnt < k ? cnt : k;
  if (kept >= k) { tau_s = s[k - 1]; tau_i = ix[k - 1]; } else { tau_s = -INFINITY; tau_i = -1; }
  return kept;
}

// grid-stride over documents, one warp per document
__global__ void __launch_bounds__(BL_WARPS * 32)
blend_scan_kernel(const float* __restrict__ q, float q_norm, const float* __restrict__ docs, int64_t N, int D,
                  const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                  const double* __restrict__ data, const int32_t* __restrict__ q_idx, const double* __restrict__ q_val,
                  int q_nnz, double alpha, int k, double* __restrict__ part_s, int* __restrict__ part_i,
                  double* __restrict__ combined_out) {
  __shared__ double buf_s[BL_WARPS][BL_CAP];
  __shared__ int buf_i[BL_WARPS][BL_CAP];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double tau_s = -INFINITY;
  int tau_i = -1;
  int cnt = 0;
  const int64_t gw = (int64_t)blockIdx.x * BL_WARPS + warp, nw = (int64_t)gridDim.x * BL_WARPS;
  for (int64_t d = gw; d < N; d += nw) {
    const float* row = docs + d * D;
    float dot = 0.f, nn = 0.f;
    for (int c = lane * 4; c < D; c += 128) {
      const float4 v = ld_stream_f4(reinterpret_cast<const float4*>(row + c));
      const float4 w = *reinterpret_cast<const float4*>(q + c);
      dot = fmaf(v.x, w.x, dot); dot = fmaf(v.y, w.y, dot); dot = fmaf(v.z, w.z, dot); dot = fmaf(v.w, w.w, dot);
      nn = fmaf(v.x, v.x, nn); nn = fmaf(v.y, v.y, nn); nn = fmaf(v.z, v.z, nn); nn = fmaf(v.w, v.w, nn);
    }
    dot = warp_sum(dot);
    nn = warp_sum(nn);
    // sklearn cosine_similarity: rows are normalised first (zero rows stay zero), then dotted
    const float dn = sqrtf(nn);
    const float cosv = (dn > 0.f && q_norm > 0.f) ? dot / (dn * q_norm) : 0.f;
    double tf = 0.0;
    if (lane == 0 && q_nnz > 0) {
      const int64_t lo = indptr[d], hi = indptr[d + 1];
      int qp = 0;
      for (int64_t e = lo; e < hi && qp < q_nnz; ++e) {
        const int32_t c = indices[e];
        while (qp < q_nnz && q_idx[qp] < c) ++qp;
        if (qp < q_nnz && q_idx[qp] == c) tf = __dadd_rn(tf, __dmul_rn(data[e], q_val[qp]));
      }
    }
    tf = __shfl_sync(0xffffffffu, tf, 0);
    // alpha * dense (float32 array times python float -> float32), then float64 sum
    const double comb = __dadd_rn((double)((float)alpha * cosv), __dmul_rn(1.0 - alpha, tf));
    if (combined_out && lane == 0) combined_out[d] = comb;
    if (blend_better(comb, (int)d, tau_s, tau_i)) {      // warp-uniform
      if (lane == 0) { buf_s[warp][cnt] = comb; buf_i[warp][cnt] = (int)d; }
      ++cnt;
      if (cnt == BL_CAP) {
        __syncwarp();
        cnt = warp_compact_blend(buf_s[warp], buf_i[warp], cnt, k, lane, tau_s, tau_i);
        __syncwarp();
      }
    }
  }
  __syncwarp();
  cnt = warp_compact_blend(buf_s[warp], buf_i[warp], cnt, k, lane, tau_s, tau_i);
  __syncwarp();
  const int64_t base = ((int64_t)blockIdx.x * BL_WARPS + warp) * k;
  for (int j = lane; j < k; j += 32) {
    part_s[base + j] = j < cnt ? buf_s[warp][j] : -INFINITY;
    part_i[base + j] = j < cnt ? buf_i[warp][j] : -1;
  }
}

// single CTA: merge P sorted lists of k
__global__ void __launch_bounds__(256)
blend_merge_kernel(const double* __restrict__ part_s, const int* __restrict__ part_i, int P, int k,
                   double* __restrict__ out_s, int64_t* __restrict__ out_i) {
  __shared__ double buf_s[BL_WARPS][BL_CAP];
  __shared__ int buf_i[BL_WARPS][BL_CAP];
  __shared__ int cnts[BL_WARPS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double tau_s = -INFINITY;
  int tau_i = -1;
  int cnt = 0;
  const int64_t total = (int64_t)P * k;
  for (int64_t t0 = (int64_t)warp * 32; t0 < total; t0 += 256) {
    const int64_t t = t0 + lane;
    double s = -INFINITY;
    int ix = -1;
    if (t < total) { s = part_s[t]; ix = part_i[t]; }
    const bool pass = ix >= 0 && blend_better(s, ix, tau_s, tau_i);
    const unsigned m = __ballot_sync(0xffffffffu, pass);
    if (m) {
      if (pass) {
        int pos = cnt + __popc(m & ((1u << lane) - 1));
        buf_s[warp][pos] = s;
        buf_i[warp][pos] = ix;
      }
      cnt += __popc(m);
      __syncwarp();
      if (cnt > BL_CAP - 32) {
        cnt = warp_compact_blend(buf_s[warp], buf_i[warp], cnt, k, lane, tau_s, tau_i);
        __syncwarp();
      }
    }
  }
  cnt = warp_compact_blend(buf_s[warp], buf_i[warp], cnt, k, lane, tau_s, tau_i);
  if (lane == 0) cnts[warp] = cnt;
  __syncthreads();
  // final: warp 0 merges the 8 lists of <= k through its own buffer
  if (warp == 0) {
    double ts = -INFINITY;
    int ti = -1;
    int c0 = cnts[0];
    for (int w = 1; w < BL_WARPS; ++w) {
      for (int j = lane; j < cnts[w]; j += 32) {
        // buffer 0 holds at most k <= 64 entries after compaction: room for another 64
        buf_s[0][c0 + j] = buf_s[w][j];
        buf_i[0][c0 + j] = buf_i[w][j];
      }
      __syncwarp();
      c0 = warp_compact_blend(buf_s[0], buf_i[0], c0 + cnts[w], k, lane, ts, ti);
      __syncwarp();
    }
    for (int j = lane; j < k; j += 32) {
      out_s[j] = j < c0 ? buf_s[0][j] : -INFINITY;
      out_i[j] = j < c0 ? (int64_t)buf_i[0][j] : (int64_t)-1;
    }
  }
}

}  // namespace ttr

extern "C" int64_t ttr_blend_topk_workspace_bytes(int k) {
  const int parts = 4 * ttr::sm_count() * ttr::BL_WARPS;
  return (int64_t)parts * k * 12 + 256;
}

extern "C" int ttr_blend_topk(const float* q, float q_norm, const float* docs, int64_t N, int D,
                              const int64_t* indptr, const int32_t* indices, const double* data,
                              const int32_t* q_idx, const double* q_val, int q_nnz, double alpha, int k,
                              double* out_scores, int64_t* out_idx, double* combined_out, void* workspace,
                              void* stream) {
  using namespace ttr;
  TTR_REQUIRE(N >= 1 && N < ((int64_t)1 << 31) && D >= 4 && D % 4 == 0, "ttr_blend_topk: bad shape (N=%lld, D=%d)",
              (long long)N, D);
  TTR_REQUIRE(k >= 1 && k <= BL_KMAX, "ttr_blend_topk: k=%d outside [1, %d]", k, BL_KMAX);
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = 4 * sm_count();
  const int parts = grid * BL_WARPS;
  double* part_s = reinterpret_cast<double*>(workspace);
  int* part_i = reinterpret_cast<int*>(part_s + (int64_t)parts * k);
  blend_scan_kernel<<<grid, BL_WARPS * 32, 0, st>>>(q, q_norm, docs, N, D, indptr, indices, data, q_idx, q_val, q_nnz,
                                                  alpha, k, part_s, part_i, combined_out);
  TTR_CHECK_LAUNCH();
  blend_merge_kernel<<<1, 256, 0, st>>>(part_s, part_i, parts, k, out_scores, out_idx);
  TTR_CHECK_LAUNCH();
  return TTR_OK;
}
