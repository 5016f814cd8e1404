// optim.cu — global-norm clipping + Adam over one flat fp32 bucket, two launches.
//
// Replaces `torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=1.0)` followed by
// `optimizer.step()` with torch.optim.Adam defaults (backend/main.py:222,257-259).  All
// trainable tensors of both towers live in one contiguous buffer (the same buffer the
// data-parallel all-reduce runs on), so the norm is one reduction and the update one pass.
#include "common.cuh"

namespace ttr {

constexpr int OPT_MAX_PARTS = 1024;

__global__ void __launch_bounds__(256)
sumsq_kernel(const float* __restrict__ g, int64_t n, float scale, float* __restrict__ parts) {
  __shared__ float sm[8];
  float acc = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = g[i] * scale;
    acc = fmaf(v, v, acc);
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += sm[i];
    parts[blockIdx.x] = s;
  }
}

__global__ void __launch_bounds__(256)
clip_adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                 int64_t n, float scale, float max_norm, float step_size, float beta1, float beta2,
                 float bc2_sqrt, float eps, const float* __restrict__ parts, int nparts,
                 float* __restrict__ norm_out) {
  __shared__ float s_coef;
  if (threadIdx.x == 0) {
    // every CTA reduces the partials in the same fixed order -> identical clip coefficient
    double tot = 0.0;
    for (int i = 0; i < nparts; ++i) tot += (double)parts[i];
    const float total = (float)sqrt(tot);
    float coef = 1.0f;
    if (max_norm > 0.f) coef = fminf(max_norm / (total + 1e-6f), 1.0f);
    s_coef = coef * scale;
    if (blockIdx.x == 0 && norm_out) *norm_out = total;
  }
  __syncthreads();
  const float coef = s_coef;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gr = g[i] * coef;
    float mi = m[i], vi = v[i];
    mi = mi + (gr - mi) * (1.0f - beta1);                 // exp_avg.lerp_(grad, 1 - beta1)
    vi = vi * beta2 + (1.0f - beta2) * gr * gr;           // exp_avg_sq.mul_(b2).addcmul_(g, g, 1 - b2)
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = p[i] - step_size * (mi / denom);               // param.addcdiv_(exp_avg, denom, -step_size)
    m[i] = mi;
    v[i] = vi;
  }
}

}  // namespace ttr

extern "C" int ttr_clip_adam(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                             float grad_scale, float max_norm, float lr, float beta1, float beta2, float eps,
                             int step, float* norm_out, float* workspace, void* stream) {
  using namespace ttr;
  TTR_REQUIRE(n >= 1 && step >= 1, "ttr_clip_adam: bad arguments (n=%lld, step=%d)", (long long)n, step);
  TTR_REQUIRE(workspace != nullptr, "ttr_clip_adam: workspace of 1024 floats required");
  cudaStream_t st = (cudaStream_t)stream;
  int parts = (int)std::min<int64_t>(OPT_MAX_PARTS, std::max<int64_t>(1, ceil_div64(n, 256 * 8)));
  sumsq_kernel<<<parts, 256, 0, st>>>(grads, n, grad_scale, workspace);
  TTR_CHECK_LAUNCH();
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  const float step_size = (float)((double)lr / bc1);
  const float bc2_sqrt = (float)sqrt(bc2);
  int grid = (int)std::min<int64_t>(4 * sm_count(), ceil_div64(n, 256));
  clip_adam_kernel<<<grid, 256, 0, st>>>(params, grads, exp_avg, exp_avg_sq, n, grad_scale, max_norm, step_size,
                                         beta1, beta2, bc2_sqrt, eps, workspace, parts, norm_out);
  TTR_CHECK_LAUNCH();
  return TTR_OK;
}
