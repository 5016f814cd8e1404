// seq_plan.cu — lengths, length-sorted packing plan and the embedding-row gather.
//
// Replaces the host-synchronising `lengths = (x != 0).sum(1).cpu()` +
// `pack_padded_sequence(enforce_sorted=False)` + `nn.Embedding` lookups of the reference
// (backend/model.py:49-57).  The packed layout is batch-major in length-sorted order:
// sorted position s holds row order[s]; its tokens occupy rows offsets[s] .. offsets[s+1]
// of every per-token matrix (X, gi, y, ...).  Lengths are non-increasing in s so a tile of
// consecutive positions has similar lengths and the active rows at step t are a prefix.
#include <cuda_fp16.h>

#include "common.cuh"

namespace ttr {

// one warp per row: count non-zero ids (quirk #1: the count, not the positions)
__global__ void lengths_kernel(const int64_t* __restrict__ ids, int B, int T,
                               int32_t* __restrict__ lengths) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (warp >= B) return;
  const int64_t* row = ids + (int64_t)warp * T;
  int c = 0;
  for (int t = lane; t < T; t += 32) c += (row[t] != 0);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if (lane == 0) lengths[warp] = c;
}

// single CTA: counting sort by length (descending) + packed offsets.
// smem: hist[T+1], start[T+1] (ints), tokbase[T+1]
__global__ void __launch_bounds__(1024) plan_kernel(const int32_t* __restrict__ lengths, int B, int T,
                                                    int32_t* __restrict__ order,
                                                    int32_t* __restrict__ offsets,
                                                    int32_t* __restrict__ status) {
  extern __shared__ int32_t sm[];
  int32_t* hist = sm;                 // bin b <-> length T - b
  int32_t* start = sm + (T + 1);      // first sorted position of bin
  int32_t* tokbase = sm + 2 * (T + 1);
  const int nb = T + 1;
  for (int i = threadIdx.x; i < nb; i += blockDim.x) hist[i] = 0;
  __syncthreads();
  for (int r = threadIdx.x; r < B; r += blockDim.x) atomicAdd(&hist[T - lengths[r]], 1);
  __syncthreads();
  if (threadIdx.x == 0) {
    // serial scan over <= 8193 bins; the plan kernel is off the critical path
    int pos = 0, tok = 0, maxlen = 0;
    for (int b = 0; b < nb; ++b) {
      int c = hist[b];
      start[b] = pos;
      tokbase[b] = tok;
      if (c > 0 && maxlen == 0) maxlen = T - b;
      pos += c;
      tok += c * (T - b);
    }
    offsets[B] = tok;
    status[0] = hist[T];   // zero-length rows
    status[1] = tok;
    status[2] = maxlen;
    status[3] = 0;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nb; i += blockDim.x) hist[i] = 0;   // reuse as cursor
  __syncthreads();
  // Deterministic placement inside a bin would need a stable pass; rows of equal length
  // are interchangeable for every consumer, so the atomic order is fine.
  for (int r = threadIdx.x; r < B; r += blockDim.x) {
    int len = lengths[r];
    int b = T - len;
    int k = atomicAdd(&hist[b], 1);
    int s = start[b] + k;
    order[s] = r;
    offsets[s] = tokbase[b] + k * len;
  }
}

// one CTA per sorted position, warps stride over timesteps, lanes over float4 columns
__global__ void __launch_bounds__(128) gather_kernel(const int64_t* __restrict__ ids, int T,
                                                     const float* __restrict__ table, int64_t V, int E,
                                                     const int32_t* __restrict__ order,
                                                     const int32_t* __restrict__ offsets,
                                                     float* __restrict__ X, int round) {
  const int s = blockIdx.x;
  const int row = order[s];
  const int off = offsets[s];
  const int len = offsets[s + 1] - off;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t* idrow = ids + (int64_t)row * T;
  const int E4 = E >> 2;
  for (int t = warp; t < len; t += 4) {
    int64_t id = idrow[t];
    if (id < 0 || id >= V) id = 0;   // torch would raise; keep memory-safe
    const float* src = table + id * E;
    if (round == 2) {
      // fp16 rows for the kind::f16 projection (E % 4 == 0): 8-byte stores of 4 halves
      uint2* d2 = reinterpret_cast<uint2*>(reinterpret_cast<__half*>(X) + (int64_t)(off + t) * E);
      const float4* s4 = reinterpret_cast<const float4*>(src);
      for (int c = lane; c < E4; c += 32) {
        const float4 v = __ldg(s4 + c);
        const __half2 lo = __floats2half2_rn(v.x, v.y), hi = __floats2half2_rn(v.z, v.w);
        d2[c] = make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
      }
      continue;
    }
    float* dst = X + (int64_t)(off + t) * E;
    if ((E & 3) == 0) {
      const float4* s4 = reinterpret_cast<const float4*>(src);
      float4* d4 = reinterpret_cast<float4*>(dst);
      for (int c = lane; c < E4; c += 32) {
        float4 v = __ldg(s4 + c);
        if (round) { v.x = round_tf32(v.x); v.y = round_tf32(v.y); v.z = round_tf32(v.z); v.w = round_tf32(v.w); }
        d4[c] = v;
      }
    } else {
      for (int c = lane; c < E; c += 32) {
        float v = __ldg(src + c);
        dst[c] = round ? round_tf32(v) : v;
      }
    }
  }
}

// dtable[id] += dX[token] for every packed token; id 0 is nn.Embedding's padding_idx
// (backend/model.py:24) and receives no gradient.
__global__ void __launch_bounds__(128) scatter_grad_kernel(const int64_t* __restrict__ ids, int T, int64_t V, int E,
                                                          const int32_t* __restrict__ order,
                                                          const int32_t* __restrict__ offsets,
                                                          const float* __restrict__ dX, float* __restrict__ dtable) {
  const int s = blockIdx.x;
  const int row = order[s];
  const int off = offsets[s];
  const int len = offsets[s + 1] - off;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int t = warp; t < len; t += 4) {
    const int64_t id = ids[(int64_t)row * T + t];
    if (id <= 0 || id >= V) continue;
    const float* src = dX + (int64_t)(off + t) * E;
    float* dst = dtable + id * E;
    for (int c = lane; c < E; c += 32) atomicAdd(dst + c, src[c]);
  }
}

}  // namespace ttr

extern "C" int ttr_embed_scatter_grad(const int64_t* ids, int B, int T, int64_t V, int E, const int32_t* order,
                                      const int32_t* offsets, const float* dX, float* dtable, void* stream) {
  using namespace ttr;
  TTR_REQUIRE(B > 0 && T > 0 && E > 0, "ttr_embed_scatter_grad: bad shape");
  scatter_grad_kernel<<<B, 128, 0, (cudaStream_t)stream>>>(ids, T, V, E, order, offsets, dX, dtable);
  TTR_CHECK_LAUNCH();
  return TTR_OK;
}

extern "C" int ttr_seq_plan(const int64_t* ids, int B, int T, int32_t* lengths, int32_t* order,
                            int32_t* offsets, int32_t* status, void* stream) {
  using namespace ttr;
  TTR_REQUIRE(B > 0 && T > 0, "ttr_seq_plan: empty batch (B=%d, T=%d)", B, T);
  TTR_REQUIRE(T <= 8192, "ttr_seq_plan: T=%d exceeds 8192", T);
  TTR_REQUIRE((int64_t)B * T < (int64_t)1 << 31, "ttr_seq_plan: B*T overflows int32 token offsets");
  cudaStream_t st = (cudaStream_t)stream;
  int threads = 256;
  int blocks = ceil_div(B * 32, threads);
  lengths_kernel<<<blocks, threads, 0, st>>>(ids, B, T, lengths);
  TTR_CHECK_LAUNCH();
  size_t smem = 3 * (size_t)(T + 1) * sizeof(int32_t);
  if (smem > 48 * 1024)
    TTR_CHECK_CUDA(cudaFuncSetAttribute(plan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  plan_kernel<<<1, 1024, smem, st>>>(lengths, B, T, order, offsets, status);
  TTR_CHECK_LAUNCH();
  return TTR_OK;
}

namespace ttr {
__global__ void f32_to_f16_kernel(const float* __restrict__ src, __half* __restrict__ dst, int64_t n) {
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 2;
  if (i + 1 < n) {
    const float2 v = *reinterpret_cast<const float2*>(src + i);
    *reinterpret_cast<__half2*>(dst + i) = __floats2half2_rn(v.x, v.y);
  } else if (i < n) {
    dst[i] = __float2half_rn(src[i]);
  }
}
}  // namespace ttr

extern "C" int ttr_f32_to_f16(const float* src, void* dst, int64_t n, void* stream) {
  using namespace ttr;
  TTR_REQUIRE(n >= 1 && ((uintptr_t)src & 7) == 0 && ((uintptr_t)dst & 3) == 0, "ttr_f32_to_f16: bad arguments");
  const int64_t pairs = (n + 1) / 2;
  f32_to_f16_kernel<<<(unsigned)((pairs + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, reinterpret_cast<__half*>(dst), n);
  TTR_CHECK_LAUNCH();
  return TTR_OK;
}

extern "C" int ttr_embed_gather(const int64_t* ids, int B, int T, const float* table, int64_t V, int E,
                                const int32_t* order, const int32_t* offsets, float* X, int round_tf32,
                                void* stream) {
  using namespace ttr;
  TTR_REQUIRE(B > 0 && T > 0 && E > 0, "ttr_embed_gather: bad shape");
  TTR_REQUIRE(round_tf32 != 2 || (E & 3) == 0, "ttr_embed_gather: fp16 output needs E %% 4 == 0 (E=%d)", E);
  gather_kernel<<<B, 128, 0, (cudaStream_t)stream>>>(ids, T, table, V, E, order, offsets, X, round_tf32);
  TTR_CHECK_LAUNCH();
  return TTR_OK;
}
