// head.cu — tower head (projection + L2 normalise), cosine triplet loss, batch metrics,
// and the small reductions the backward pass needs.
//
// Replaces `cat(h_n[-2], h_n[-1])` -> `self.projection` -> `F.normalize` (backend/model.py:65-74),
// `triplet_loss_cosine` (backend/model.py:109-114) and `compute_batch_metrics`
// (backend/trainer.py:38-55).
#include "common.cuh"

namespace ttr {

constexpr int PR = 8;   // rows per CTA in the head kernel

// out[b, j] = sum_k x[b, k] W[j, k] + bias[j]; optional L2 normalise.  8 rows per CTA are
// staged in shared memory; a warp owns output units j = warp, warp+8, ... and streams W rows
// with coalesced loads (W is 512 KB and stays in L2 across CTAs).
__global__ void __launch_bounds__(256)
proj_l2norm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                       int B, int IN, int H, int normalize, float* __restrict__ out, float* __restrict__ raw_out) {
  extern __shared__ __align__(16) float sm[];
  float* xs = sm;             // [PR][IN]
  float* ys = sm + PR * IN;   // [PR][H]
  const int b0 = blockIdx.x * PR;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < PR * IN; i += blockDim.x) {
    int r = i / IN, c = i % IN;
    xs[i] = (b0 + r < B) ? x[(int64_t)(b0 + r) * IN + c] : 0.f;
  }
  __syncthreads();
  if (w) {
    for (int j = warp; j < H; j += 8) {
      float acc[PR];
#pragma unroll
      for (int r = 0; r < PR; ++r) acc[r] = 0.f;
      const float* wr = w + (int64_t)j * IN;
      for (int k = lane; k < IN; k += 32) {
        const float wv = __ldg(wr + k);
#pragma unroll
        for (int r = 0; r < PR; ++r) acc[r] = fmaf(xs[r * IN + k], wv, acc[r]);
      }
#pragma unroll
      for (int r = 0; r < PR; ++r) acc[r] = warp_sum(acc[r]);
      if (lane < PR) {
        float v = acc[0];
#pragma unroll
        for (int r = 1; r < PR; ++r) v = (lane == r) ? acc[r] : v;
        ys[lane * H + j] = v + (bias ? bias[j] : 0.f);
      }
    }
  } else {
    for (int i = threadIdx.x; i < PR * H; i += blockDim.x) ys[i] = xs[(i / H) * IN + (i % H)];
  }
  __syncthreads();
  // warp r finishes row r
  const int r = warp;
  if (b0 + r < B) {
    float ss = 0.f;
    for (int j = lane; j < H; j += 32) ss = fmaf(ys[r * H + j], ys[r * H + j], ss);
    ss = warp_sum(ss);
    const float inv = normalize ? 1.0f / fmaxf(sqrtf(ss), 1e-12f) : 1.0f;
    for (int j = lane; j < H; j += 32) {
      const float v = ys[r * H + j];
      if (raw_out) raw_out[(int64_t)(b0 + r) * H + j] = v;
      out[(int64_t)(b0 + r) * H + j] = normalize ? v * inv : v;
    }
  }
}

// d_raw = (d_out - y (y . d_out)) / max(||raw||, eps), y = raw / max(||raw||, eps)
__global__ void l2norm_bwd_kernel(const float* __restrict__ d_out, const float* __restrict__ raw, int B, int H,
                                  int normalize, float* __restrict__ d_raw) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= B) return;
  const float* v = raw + (int64_t)row * H;
  const float* g = d_out + (int64_t)row * H;
  float* o = d_raw + (int64_t)row * H;
  if (!normalize) {
    for (int j = lane; j < H; j += 32) o[j] = g[j];
    return;
  }
  float ss = 0.f, dot = 0.f;
  for (int j = lane; j < H; j += 32) { ss = fmaf(v[j], v[j], ss); dot = fmaf(v[j], g[j], dot); }
  ss = warp_sum(ss); dot = warp_sum(dot);
  const float nrm = sqrtf(ss);
  if (nrm > 1e-12f) {
    const float inv = 1.0f / nrm;
    const float c = dot * inv * inv * inv;        // (y.g)/||v|| with y = v/||v||
    for (int j = lane; j < H; j += 32) o[j] = g[j] * inv - v[j] * c;
  } else {
    for (int j = lane; j < H; j += 32) o[j] = g[j] * 1e12f;
  }
}

struct Cos3 { float qp, qn, nq, np, nn; };

__device__ __forceinline__ Cos3 row_cos(const float* q, const float* p, const float* n, int H, int lane) {
  float qq = 0.f, pp = 0.f, nn = 0.f, qp = 0.f, qn = 0.f;
  for (int j = lane; j < H; j += 32) {
    const float a = q[j], b = p[j], c = n[j];
    qq = fmaf(a, a, qq); pp = fmaf(b, b, pp); nn = fmaf(c, c, nn);
    qp = fmaf(a, b, qp); qn = fmaf(a, c, qn);
  }
  qq = warp_sum(qq); pp = warp_sum(pp); nn = warp_sum(nn); qp = warp_sum(qp); qn = warp_sum(qn);
  Cos3 r;
  r.nq = fmaxf(sqrtf(qq), 1e-8f); r.np = fmaxf(sqrtf(pp), 1e-8f); r.nn = fmaxf(sqrtf(nn), 1e-8f);
  r.qp = qp / (r.nq * r.np);
  r.qn = qn / (r.nq * r.nn);
  return r;
}

__global__ void __launch_bounds__(256)
triplet_fwd_kernel(const float* __restrict__ q, const float* __restrict__ p, const float* __restrict__ n, int B,
                   int H, float margin, float* __restrict__ loss_out) {
  __shared__ float part[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float acc = 0.f;
  for (int row = blockIdx.x * 8 + warp; row < B; row += gridDim.x * 8) {
    Cos3 c = row_cos(q + (int64_t)row * H, p + (int64_t)row * H, n + (int64_t)row * H, H, lane);
    acc += fmaxf(c.qn - c.qp + margin, 0.f);
  }
  if (lane == 0) part[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += part[i];
    atomicAdd(loss_out, s / (float)B);
  }
}

__global__ void __launch_bounds__(256)
triplet_bwd_kernel(const float* __restrict__ q, const float* __restrict__ p, const float* __restrict__ n, int B,
                   int H, float margin, const float* __restrict__ dloss, float* __restrict__ dq,
                   float* __restrict__ dp, float* __restrict__ dn) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= B) return;
  const float* qr = q + (int64_t)row * H;
  const float* pr = p + (int64_t)row * H;
  const float* nr = n + (int64_t)row * H;
  Cos3 c = row_cos(qr, pr, nr, H, lane);
  const float pre = c.qn - c.qp + margin;
  const float g = (pre >= 0.f ? 1.0f : 0.0f) * (dloss ? *dloss : 1.0f) / (float)B;   // clamp(min=0) passes at 0
  // d cos(a,b)/da = (b/|b| - cos * a/|a|) / |a|
  const float iq = 1.0f / c.nq, ip = 1.0f / c.np, in_ = 1.0f / c.nn;
  for (int j = lane; j < H; j += 32) {
    const float a = qr[j] * iq, b = pr[j] * ip, d = nr[j] * in_;
    const float dcos_qn_dq = (d - c.qn * a) * iq;
    const float dcos_qp_dq = (b - c.qp * a) * iq;
    dq[(int64_t)row * H + j] = g * (dcos_qn_dq - dcos_qp_dq);
    dp[(int64_t)row * H + j] = -g * (a - c.qp * b) * ip;
    dn[(int64_t)row * H + j] = g * (a - c.qn * d) * in_;
  }
}

__global__ void __launch_bounds__(256)
batch_metrics_kernel(const float* __restrict__ q, const float* __restrict__ p, const float* __restrict__ n,
                     int B, int H, float* __restrict__ out) {
  __shared__ float part[8][5];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  for (int row = blockIdx.x * 8 + warp; row < B; row += gridDim.x * 8) {
    const float* qr = q + (int64_t)row * H;
    const float* pr = p + (int64_t)row * H;
    const float* nr = n + (int64_t)row * H;
    float qq = 0.f, qp = 0.f, qn = 0.f;
    for (int j = lane; j < H; j += 32) {
      qq = fmaf(qr[j], qr[j], qq); qp = fmaf(qr[j], pr[j], qp); qn = fmaf(qr[j], nr[j], qn);
    }
    qq = warp_sum(qq); qp = warp_sum(qp); qn = warp_sum(qn);
    acc[0] += (qp > qn) ? 1.f : 0.f;
    acc[1] += qp - qn;
    acc[2] += sqrtf(qq);
    acc[3] += qp;
    acc[4] += qn;
  }
  if (lane == 0)
    for (int i = 0; i < 5; ++i) part[warp][i] = acc[i];
  __syncthreads();
  if (threadIdx.x < 5) {
    float s = 0.f;
    for (int wv = 0; wv < 8; ++wv) s += part[wv][threadIdx.x];
    atomicAdd(out + threadIdx.x, s / (float)B);
  }
}

// out[j] (+)= sum_i A[i, j] over valid rows.  grid (col blocks of 128, row splits)
__global__ void __launch_bounds__(128)
colsum_kernel(const float* __restrict__ A, int m_bound, const int32_t* __restrict__ m_valid, int N,
              float* __restrict__ out) {
  const int M = m_valid ? min(m_bound, *m_valid) : m_bound;
  const int j = blockIdx.x * 128 + threadIdx.x;
  if (j >= N) return;
  const int per = ceil_div(M, (int)gridDim.y);
  const int i0 = blockIdx.y * per, i1 = min(M, i0 + per);
  float s = 0.f;
  for (int i = i0; i < i1; ++i) s += A[(int64_t)i * N + j];
  if (i1 > i0) atomicAdd(out + j, s);
}

}  // namespace ttr

extern "C" int ttr_proj_l2norm_fwd(const float* h_cat, const float* w, const float* bias, int B, int IN, int H,
                                   int normalize, float* out, float* raw_out, void* stream) {
  using namespace ttr;
  TTR_REQUIRE(B >= 1 && IN >= 1 && H >= 1, "ttr_proj_l2norm_fwd: bad shape");
  TTR_REQUIRE(w != nullptr || IN == H, "ttr_proj_l2norm_fwd: no projection requires IN == H");
  const size_t smem = (size_t)PR * (IN + H) * sizeof(float);
  TTR_REQUIRE(smem <= 200 * 1024, "ttr_proj_l2norm_fwd: IN+H too large");
  if (smem > 48 * 1024)
    TTR_CHECK_CUDA(cudaFuncSetAttribute(proj_l2norm_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  proj_l2norm_fwd_kernel<<<ceil_div(B, PR), 256, smem, (cudaStream_t)stream>>>(h_cat, w, bias, B, IN, H, normalize,
                                                                              out, raw_out);
  TTR_CHECK_LAUNCH();
  return TTR_OK;
}

extern "C" int ttr_l2norm_bwd(const float* d_out, const float* raw, int B, int H, int normalize, float* d_raw,
                              void* stream) {
  using namespace ttr;
  TTR_REQUIRE(B >= 1 && H >= 1, "ttr_l2norm_bwd: bad shape");
  l2norm_bwd_kernel<<<ceil_div(B, 8), 256, 0, (cudaStream_t)stream>>>(d_out, raw, B, H, normalize, d_raw);
  TTR_CHECK_LAUNCH();
  return TTR_OK;
}

extern "C" int ttr_triplet_fwd(const float* q, const float* p, const float* n, int B, int H, float margin,
                               float* loss_out, void* stream) {
  using namespace ttr;
  TTR_REQUIRE(B >= 1 && H >= 1, "ttr_triplet_fwd: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  TTR_CHECK_CUDA(cudaMemsetAsync(loss_out, 0, sizeof(float), st));
  int grid = min(ceil_div(B, 8), 4 * sm_count());
  triplet_fwd_kernel<<<grid, 256, 0, st>>>(q, p, n, B, H, margin, loss_out);
  TTR_CHECK_LAUNCH();
  return TTR_OK;
}

extern "C" int ttr_triplet_bwd(const float* q, const float* p, const float* n, int B, int H, float margin,
                               const float* dloss, float* dq, float* dp, float* dn, void* stream) {
  using namespace ttr;
  TTR_REQUIRE(B >= 1 && H >= 1, "ttr_triplet_bwd: bad shape");
  triplet_bwd_kernel<<<ceil_div(B, 8), 256, 0, (cudaStream_t)stream>>>(q, p, n, B, H, margin, dloss, dq, dp, dn);
  TTR_CHECK_LAUNCH();
  return TTR_OK;
}

extern "C" int ttr_batch_metrics(const float* q, const float* p, const float* n, int B, int H, float* out,
                                 void* stream) {
  using namespace ttr;
  TTR_REQUIRE(B >= 1 && H >= 1, "ttr_batch_metrics: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  TTR_CHECK_CUDA(cudaMemsetAsync(out, 0, 5 * sizeof(float), st));
  int grid = min(ceil_div(B, 8), 4 * sm_count());
  batch_metrics_kernel<<<grid, 256, 0, st>>>(q, p, n, B, H, out);
  TTR_CHECK_LAUNCH();
  return TTR_OK;
}

extern "C" int ttr_colsum(const float* A, int m_bound, const int32_t* m_valid, int N, float* out, int accumulate,
                          void* stream) {
  using namespace ttr;
  TTR_REQUIRE(m_bound >= 1 && N >= 1, "ttr_colsum: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  if (!accumulate) TTR_CHECK_CUDA(cudaMemsetAsync(out, 0, (size_t)N * sizeof(float), st));
  int splits = max(1, min(ceil_div(m_bound, 256), (4 * sm_count()) / max(1, ceil_div(N, 128))));
  dim3 grid(ceil_div(N, 128), splits);
  colsum_kernel<<<grid, 128, 0, st>>>(A, m_bound, m_valid, N, out);
  TTR_CHECK_LAUNCH();
  return TTR_OK;
}
