// gru_fwd.cu — persistent GRU recurrence (forward), one layer, all directions.
//
// Replaces the sequential half of `self.rnn(packed)` (backend/model.py:59-62).
//
// H == 256 kernel: a thread-block cluster of 8 CTAs owns one tile of 64 length-sorted rows
// and one direction for ALL timesteps.  CTA c owns hidden units [32c, 32c+32); each of its
// 256 threads keeps the 3x32 slice W_h{r,z,n}[unit, 32*kc .. 32*kc+31] of W_hh in REGISTERS
// for the whole kernel (W_hh never touches shared memory or HBM again).  Per step a lane
// accumulates 8 rows x 3 gates over its k-chunk from the shared-memory copy of h, the 8
// k-chunks are combined with a halving warp-shuffle butterfly that leaves lane (unit, kc)
// holding the three gate pre-activations of row 8*rg+kc, the gate math is fused, and the
// new h value is pushed straight into all 8 CTAs' next-step buffers through distributed
// shared memory.  One cluster barrier per step.
//
// Other H: a generic one-CTA-per-tile kernel that streams W_hh from L2 (correct for any H;
// used by the small-shape parity tests and non-default configs).
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace ttr {

constexpr int GH = 256;
constexpr int GCL = 8;                 // CTAs per cluster
constexpr int GUN = GH / GCL;          // 32 hidden units per CTA
constexpr int GBT = 64;                // rows per tile
constexpr int GCS = GBT * 32 + 4;      // chunk stride (floats): +4 keeps the 8 k-chunks on distinct banks
constexpr int GTHREADS = 256;

int g_debug_flags = 0;

struct GruFwdArgs {
  const float* gi;
  const float* w_hh;
  const float* b_hh;
  const int32_t* order;
  const int32_t* offsets;
  int B, dirs;
  float* y;
  float* h_last;
  float* saved;
};

__global__ void __cluster_dims__(GCL, 1, 1) __launch_bounds__(GTHREADS, 1)
gru_fwd_cluster_kernel(GruFwdArgs a) {
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ __align__(16) float smem[];
  float* hbuf = smem;                                   // [2][GCL][GCS]
  int* lens = reinterpret_cast<int*>(smem + 2 * GCL * GCS);   // [GBT]
  int* toff = lens + GBT;
  int* rowid = toff + GBT;

  const int rank = (int)cluster.block_rank();
  const int tile = blockIdx.x / GCL;
  const int dir = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int usub = lane >> 3, kc = lane & 7;
  const int u = warp * 4 + usub;          // unit inside this CTA's slice
  const int j = rank * GUN + u;           // global hidden unit
  const int G3 = 3 * GH;
  const int s0 = tile * GBT;

  for (int i = threadIdx.x; i < GBT; i += GTHREADS) {
    int s = s0 + i;
    if (s < a.B) {
      int off = a.offsets[s];
      lens[i] = a.offsets[s + 1] - off;
      toff[i] = off;
      rowid[i] = a.order[s];
    } else {
      lens[i] = 0; toff[i] = 0; rowid[i] = 0;
    }
  }
  for (int i = threadIdx.x; i < 2 * GCL * GCS; i += GTHREADS) hbuf[i] = 0.f;

  // resident weights: W_hh[dir][g*H + j][kc*32 .. +31]
  float w[3][32];
  {
    const float* wbase = a.w_hh + (size_t)dir * G3 * GH;
#pragma unroll
    for (int g = 0; g < 3; ++g) {
      const float4* src = reinterpret_cast<const float4*>(wbase + (size_t)(g * GH + j) * GH + kc * 32);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float4 v = __ldg(src + i);
        w[g][4 * i + 0] = v.x; w[g][4 * i + 1] = v.y; w[g][4 * i + 2] = v.z; w[g][4 * i + 3] = v.w;
      }
    }
  }
  const float bh_r = a.b_hh[dir * G3 + j];
  const float bh_z = a.b_hh[dir * G3 + GH + j];
  const float bh_n = a.b_hh[dir * G3 + 2 * GH + j];

  float* remote[GCL];
#pragma unroll
  for (int c = 0; c < GCL; ++c) remote[c] = cluster.map_shared_rank(hbuf, c);

  __syncthreads();
  cluster.sync();

  const int maxlen = lens[0];
  const int gi_ld = a.dirs * G3;
  const int y_ld = a.dirs * GH;

  for (int t = 0; t < maxlen; ++t) {
    const int cur = t & 1, nxt = cur ^ 1;
    int cnt = (lens[lane] > t) + (lens[lane + 32] > t);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    const int nact = cnt;
    const int ngroups = (nact + 7) >> 3;
    const float* hcur = hbuf + cur * GCL * GCS;
    const int nxt_base = nxt * GCL * GCS + rank * GCS;

    for (int rg = 0; rg < ngroups; ++rg) {
      const int row = rg * 8 + kc;          // the row this lane finalises
      const bool active = row < nact;
      int tok = 0, len = 0;
      float gir = 0.f, giz = 0.f, gin = 0.f;
      if (active) {
        len = lens[row];
        tok = toff[row] + (dir == 0 ? t : len - 1 - t);
        const float* g = a.gi + (size_t)tok * gi_ld + dir * G3 + j;
        gir = __ldg(g); giz = __ldg(g + GH); gin = __ldg(g + 2 * GH);
      }
      float acc[8][3];
#pragma unroll
      for (int r = 0; r < 8; ++r) { acc[r][0] = 0.f; acc[r][1] = 0.f; acc[r][2] = 0.f; }
      const float* hb = hcur + kc * GCS + rg * 8 * 32;
#pragma unroll
      for (int kg = 0; kg < 8; ++kg) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const float4 hv = *reinterpret_cast<const float4*>(hb + r * 32 + kg * 4);
#pragma unroll
          for (int g = 0; g < 3; ++g) {
            acc[r][g] = fmaf(hv.x, w[g][kg * 4 + 0], acc[r][g]);
            acc[r][g] = fmaf(hv.y, w[g][kg * 4 + 1], acc[r][g]);
            acc[r][g] = fmaf(hv.z, w[g][kg * 4 + 2], acc[r][g]);
            acc[r][g] = fmaf(hv.w, w[g][kg * 4 + 3], acc[r][g]);
          }
        }
      }
      // halving butterfly over the 8 k-chunks (lane bits 2,1,0): row = kc survives
      float a4[4][3], a2[2][3], a1[3];
      {
        const bool up = (lane & 4) != 0;
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int g = 0; g < 3; ++g) {
            float send = up ? acc[r][g] : acc[r + 4][g];
            float keep = up ? acc[r + 4][g] : acc[r][g];
            a4[r][g] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
          }
      }
      {
        const bool up = (lane & 2) != 0;
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
          for (int g = 0; g < 3; ++g) {
            float send = up ? a4[r][g] : a4[r + 2][g];
            float keep = up ? a4[r + 2][g] : a4[r][g];
            a2[r][g] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
          }
      }
      {
        const bool up = (lane & 1) != 0;
#pragma unroll
        for (int g = 0; g < 3; ++g) {
          float send = up ? a2[0][g] : a2[1][g];
          float keep = up ? a2[1][g] : a2[0][g];
          a1[g] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
        }
      }
      if (active) {
        const float hprev = hcur[rank * GCS + row * 32 + u];
        const float r = sigmoidf_acc(gir + a1[0] + bh_r);
        const float z = sigmoidf_acc(giz + a1[1] + bh_z);
        const float ghn = a1[2] + bh_n;
        const float n = tanhf(fmaf(r, ghn, gin));
        const float hnew = fmaf(z, hprev - n, n);      // (1-z)*n + z*h
        const int slot = nxt_base + row * 32 + u;
#pragma unroll
        for (int c = 0; c < GCL; ++c) remote[c][slot] = hnew;
        if (a.y) a.y[(size_t)tok * y_ld + dir * GH + j] = hnew;
        if (a.saved) {
          float* sv = a.saved + ((size_t)tok * a.dirs + dir) * 4 * GH + j;
          sv[0] = r; sv[GH] = z; sv[2 * GH] = n; sv[3 * GH] = ghn;
        }
        if (t == len - 1) a.h_last[(size_t)rowid[row] * y_ld + dir * GH + j] = hnew;
      }
    }
    cluster.sync();
  }
}

// ---- generic kernel: any H, 8 rows per CTA, W_hh streamed from L2 -----------------------
constexpr int GEN_ROWS = 8;

__global__ void __launch_bounds__(256)
gru_fwd_generic_kernel(GruFwdArgs a, int H) {
  extern __shared__ __align__(16) float smem[];
  float* hbuf = smem;                       // [2][GEN_ROWS][H]
  __shared__ int lens[GEN_ROWS], toff[GEN_ROWS], rowid[GEN_ROWS];
  const int dir = blockIdx.y;
  const int s0 = blockIdx.x * GEN_ROWS;
  const int G3 = 3 * H;
  if (threadIdx.x < GEN_ROWS) {
    int s = s0 + threadIdx.x;
    if (s < a.B) {
      int off = a.offsets[s];
      lens[threadIdx.x] = a.offsets[s + 1] - off;
      toff[threadIdx.x] = off;
      rowid[threadIdx.x] = a.order[s];
    } else {
      lens[threadIdx.x] = 0; toff[threadIdx.x] = 0; rowid[threadIdx.x] = 0;
    }
  }
  for (int i = threadIdx.x; i < 2 * GEN_ROWS * H; i += blockDim.x) hbuf[i] = 0.f;
  __syncthreads();
  const int maxlen = lens[0];
  const float* W = a.w_hh + (size_t)dir * G3 * H;
  const float* bh = a.b_hh + dir * G3;
  const int gi_ld = a.dirs * G3, y_ld = a.dirs * H;
  for (int t = 0; t < maxlen; ++t) {
    const float* hc = hbuf + (t & 1) * GEN_ROWS * H;
    float* hn = hbuf + ((t & 1) ^ 1) * GEN_ROWS * H;
    for (int o = threadIdx.x; o < GEN_ROWS * H; o += blockDim.x) {
      const int row = o / H, j = o % H;
      const int len = lens[row];
      if (t >= len) continue;
      const int tok = toff[row] + (dir == 0 ? t : len - 1 - t);
      const float* hr = hc + row * H;
      float ar = 0.f, az = 0.f, an = 0.f;
      const float* wr = W + (size_t)j * H;
      const float* wz = W + (size_t)(H + j) * H;
      const float* wn = W + (size_t)(2 * H + j) * H;
      for (int k = 0; k < H; ++k) {
        const float hv = hr[k];
        ar = fmaf(hv, __ldg(wr + k), ar);
        az = fmaf(hv, __ldg(wz + k), az);
        an = fmaf(hv, __ldg(wn + k), an);
      }
      const float* g = a.gi + (size_t)tok * gi_ld + dir * G3 + j;
      const float r = sigmoidf_acc(g[0] + ar + bh[j]);
      const float z = sigmoidf_acc(g[H] + az + bh[H + j]);
      const float ghn = an + bh[2 * H + j];
      const float n = tanhf(fmaf(r, ghn, g[2 * H]));
      const float hprev = hr[j];
      const float hnew = fmaf(z, hprev - n, n);
      hn[row * H + j] = hnew;
      if (a.y) a.y[(size_t)tok * y_ld + dir * H + j] = hnew;
      if (a.saved) {
        float* sv = a.saved + ((size_t)tok * a.dirs + dir) * 4 * H + j;
        sv[0] = r; sv[H] = z; sv[2 * H] = n; sv[3 * H] = ghn;
      }
      if (t == len - 1) a.h_last[(size_t)rowid[row] * y_ld + dir * H + j] = hnew;
    }
    __syncthreads();
  }
}

int launch_gru_fwd_tc(const float* gi, const float* w_hh, const float* b_hh, const int32_t* order,
                      const int32_t* offsets, int B, int dirs, float* y, float* h_last, float* saved, void* workspace,
                      bool half_io, cudaStream_t st);
int64_t gru_fwd_tc_workspace_bytes(int B, int dirs);

}  // namespace ttr

extern "C" int ttr_debug_set_flags(int flags) {
  ttr::g_debug_flags = flags;
  return TTR_OK;
}

static int gru_recurrence_fwd_impl(const float* gi, const float* w_hh, const float* b_hh, const int32_t* order,
                                  const int32_t* offsets, int B, int H, int dirs, float* y, float* h_last,
                                  float* saved, void* workspace, int64_t workspace_bytes, bool half_io, void* stream) {
  using namespace ttr;
  TTR_REQUIRE(B >= 1 && H >= 1 && (dirs == 1 || dirs == 2), "ttr_gru_recurrence_fwd: bad shape");
  TTR_REQUIRE(h_last != nullptr, "ttr_gru_recurrence_fwd: h_last is required");
  cudaStream_t st = (cudaStream_t)stream;
  GruFwdArgs a{gi, w_hh, b_hh, order, offsets, B, dirs, y, h_last, saved};
  if (H == GH && workspace != nullptr && !(g_debug_flags & (1 | 1024))) {
    // default for H = 256: recurrent product on the tensor cores (gru_fwd_tc.cu)
    TTR_REQUIRE(workspace_bytes >= gru_fwd_tc_workspace_bytes(B, dirs), "ttr_gru_recurrence_fwd_ws: workspace of %lld B < %lld B",
                (long long)workspace_bytes, (long long)gru_fwd_tc_workspace_bytes(B, dirs));
    return launch_gru_fwd_tc(gi, w_hh, b_hh, order, offsets, B, dirs, y, h_last, saved, workspace, half_io, st);
  }
  TTR_REQUIRE(!half_io, "ttr_gru_recurrence_fwd_f16: only the tcgen05 H = 256 kernel reads fp16 gi (H=%d, flags=%d)", H,
              g_debug_flags);
  if (H == GH && !(g_debug_flags & 1)) {
    // no workspace (or debug bit 10): the fp32 CUDA-core cluster kernel (W_hh in registers, warp-shuffle reductions)
    const size_t smem = (size_t)2 * GCL * GCS * sizeof(float) + 3 * GBT * sizeof(int);
    TTR_CHECK_CUDA(cudaFuncSetAttribute(gru_fwd_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(ceil_div(B, GBT) * GCL, dirs);
    gru_fwd_cluster_kernel<<<grid, GTHREADS, smem, st>>>(a);
    TTR_CHECK_LAUNCH();
  } else {
    const size_t smem = (size_t)2 * GEN_ROWS * H * sizeof(float);
    TTR_REQUIRE(smem <= 200 * 1024, "ttr_gru_recurrence_fwd: H=%d too large for the generic kernel", H);
    if (smem > 48 * 1024)
      TTR_CHECK_CUDA(cudaFuncSetAttribute(gru_fwd_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(ceil_div(B, GEN_ROWS), dirs);
    gru_fwd_generic_kernel<<<grid, 256, smem, st>>>(a, H);
    TTR_CHECK_LAUNCH();
  }
  return TTR_OK;
}

extern "C" int ttr_gru_recurrence_fwd(const float* gi, const float* w_hh, const float* b_hh,
                                      const int32_t* order, const int32_t* offsets, int B, int H, int dirs,
                                      float* y, float* h_last, float* saved, void* stream) {
  return gru_recurrence_fwd_impl(gi, w_hh, b_hh, order, offsets, B, H, dirs, y, h_last, saved, nullptr, 0, false, stream);
}

extern "C" int64_t ttr_gru_fwd_workspace_bytes(int B, int H, int dirs) {
  return H == ttr::GH ? ttr::gru_fwd_tc_workspace_bytes(B, dirs) : 0;
}

extern "C" int ttr_gru_recurrence_fwd_ws(const float* gi, const float* w_hh, const float* b_hh,
                                         const int32_t* order, const int32_t* offsets, int B, int H, int dirs,
                                         float* y, float* h_last, float* saved, void* workspace,
                                         int64_t workspace_bytes, void* stream) {
  return gru_recurrence_fwd_impl(gi, w_hh, b_hh, order, offsets, B, H, dirs, y, h_last, saved, workspace,
                                 workspace_bytes, false, stream);
}

extern "C" int ttr_gru_recurrence_fwd_f16(const void* gi16, const float* w_hh, const float* b_hh,
                                          const int32_t* order, const int32_t* offsets, int B, int H, int dirs,
                                          void* y16, float* h_last, void* workspace, int64_t workspace_bytes,
                                          void* stream) {
  TTR_REQUIRE(workspace != nullptr, "ttr_gru_recurrence_fwd_f16: workspace is required");
  return gru_recurrence_fwd_impl(reinterpret_cast<const float*>(gi16), w_hh, b_hh, order, offsets, B, H, dirs,
                                 reinterpret_cast<float*>(y16), h_last, nullptr, workspace, workspace_bytes, true, stream);
}

extern "C" int ttr_debug_get_flags(int* out) {
  *out = ttr::g_debug_flags;
  return TTR_OK;
}
