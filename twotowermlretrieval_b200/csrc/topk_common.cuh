// topk_common.cuh — exact streaming top-k building blocks shared by the scoring kernels
// and the merge kernel.
//
// A warp owns a small candidate buffer per query in shared memory.  Scores are filtered
// against the current k-th best (tau); survivors are appended; when the buffer is nearly
// full the warp sorts it (bitonic, 128 slots), keeps the best k and tightens tau.  Every
// comparison uses one total order — higher score first, then lower index — so the result
// is exact and independent of how the stream was split across warps, CTAs or GPUs.
#pragma once
#include "common.cuh"

namespace ttr {

constexpr int TOPK_CAP = 128;   // slots per (warp, query) buffer
constexpr int TOPK_KMAX = 64;   // largest supported k (k <= CAP/2)
constexpr int IDX_PAD = 0x7fffffff;

template <typename IdxT>
struct IdxTraits;
template <>
struct IdxTraits<int32_t> {
  static __device__ __forceinline__ int32_t pad() { return 0x7fffffff; }
};
template <>
struct IdxTraits<int64_t> {
  static __device__ __forceinline__ int64_t pad() { return 0x7fffffffffffffffLL; }
};

template <typename IdxT>
__device__ __forceinline__ bool key_better(float sa, IdxT ia, float sb, IdxT ib) {
  return sa > sb || (sa == sb && ia < ib);
}

// Sort n_slots (power of two, multiple of 64) slots descending by (score, idx) with one warp.
template <typename IdxT>
__device__ __forceinline__ void warp_bitonic_desc(float* s, IdxT* ix, int n_slots, int lane) {
  for (int k = 2; k <= n_slots; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = lane; t < (n_slots >> 1); t += 32) {
        // t-th compare-exchange of this stage: insert a zero bit at position log2(j)
        int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        int p = i | j;
        bool desc = ((i & k) == 0);
        float si = s[i], sp = s[p];
        IdxT ii = ix[i], ip = ix[p];
        bool p_better = key_better<IdxT>(sp, ip, si, ii);
        if (p_better == desc) {
          s[i] = sp; s[p] = si;
          ix[i] = ip; ix[p] = ii;
        }
      }
      __syncwarp();
    }
  }
}

// Same, block-wide (all threads of the CTA call it), for the per-CTA merge.
template <typename IdxT>
__device__ __forceinline__ void block_bitonic_desc(float* s, IdxT* ix, int n_slots) {
  for (int k = 2; k <= n_slots; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < (n_slots >> 1); t += blockDim.x) {
        int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        int p = i | j;
        bool desc = ((i & k) == 0);
        float si = s[i], sp = s[p];
        IdxT ii = ix[i], ip = ix[p];
        bool p_better = key_better<IdxT>(sp, ip, si, ii);
        if (p_better == desc) {
          s[i] = sp; s[p] = si;
          ix[i] = ip; ix[p] = ii;
        }
      }
      __syncthreads();
    }
  }
}

// Compact one warp buffer: pad, sort, keep k.  Returns the new count; tau_* receive the
// k-th best key (or the "accept everything" key when fewer than k entries exist).
template <typename IdxT>
__device__ __forceinline__ int warp_compact(float* s, IdxT* ix, int cnt, int k, int lane,
                                            float& tau_s, IdxT& tau_i) {
  for (int t = cnt + lane; t < TOPK_CAP; t += 32) {
    s[t] = -INFINITY;
    ix[t] = IdxTraits<IdxT>::pad();
  }
  __syncwarp();
  warp_bitonic_desc<IdxT>(s, ix, TOPK_CAP, lane);
  int kept = cnt < k ? cnt : k;
  if (kept >= k) {
    tau_s = s[k - 1];
    tau_i = ix[k - 1];
  } else {
    tau_s = -INFINITY;
    tau_i = IdxTraits<IdxT>::pad();
  }
  return kept;
}

}  // namespace ttr
