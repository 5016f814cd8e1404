// gru_bwd.cu — backpropagation through time of one GRU layer (all directions).
//
// The autograd of `self.rnn(packed)` (backend/model.py:59-62) reached from `loss.backward()`
// (backend/main.py:254).  Mirrors gru_fwd.cu: for H == 256 a cluster of 8 CTAs owns a tile of
// 32 length-sorted rows and one direction for all timesteps; CTA c owns hidden units
// [32c, 32c+32) and keeps the COLUMN slice W_hh[:, 32c..32c+31] (768 x 32) in registers.
// Per step (in reverse step order):
//   A. gate-gradient math for the CTA's own units from the saved activations; d(gi), d(gh)
//      go to global (they feed the weight-gradient GEMMs) and d(gh) is pushed into all 8
//      CTAs' shared memory (distributed shared memory);
//   B. dh_{t-1} = dh_t * z + d(gh) W_hh for the CTA's own units (register-resident weights,
//      halving shuffle butterfly across the 8 k-chunks).
// dh itself never leaves registers.  Other H: generic one-CTA-per-tile kernel.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace ttr {

extern int g_debug_flags;

constexpr int BH = 256;
constexpr int BCL = 8;
constexpr int BUN = BH / BCL;        // 32 units per CTA
constexpr int BBT = 32;              // rows per tile
constexpr int BCH = 96;              // gate rows per k-chunk (768 / 8)
constexpr int BCHP = BCH + 4;        // padded chunk stride -> conflict-free 128-bit reads
constexpr int BROW = 8 * BCHP;       // floats per row in the d(gh) buffer
constexpr int BTHREADS = 256;

struct GruBwdArgs {
  const float* dy;        // [Mtok, dirs*H] or null
  const float* dh_last;   // [B, dirs*H] or null
  const float* y;         // [Mtok, dirs*H]
  const float* saved;     // [Mtok, dirs, 4, H]
  const float* w_hh;      // [dirs, 3H, H]
  const int32_t* order;
  const int32_t* offsets;
  int B, dirs;
  float* dgi;             // [Mtok, dirs*3H]
  float* dgh;             // [Mtok, dirs*3H]
};

__global__ void __cluster_dims__(BCL, 1, 1) __launch_bounds__(BTHREADS, 1)
gru_bwd_cluster_kernel(GruBwdArgs a) {
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ __align__(16) float smem[];
  float* gbuf = smem;                                        // [BBT][BROW]
  int* lens = reinterpret_cast<int*>(smem + BBT * BROW);
  int* toff = lens + BBT;
  int* rowid = toff + BBT;

  const int rank = (int)cluster.block_rank();
  const int tile = blockIdx.x / BCL;
  const int dir = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int usub = lane >> 3, kc = lane & 7;
  const int u = warp * 4 + usub;
  const int j = rank * BUN + u;
  const int G3 = 3 * BH;
  const int s0 = tile * BBT;

  for (int i = threadIdx.x; i < BBT; i += BTHREADS) {
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
  // column slice of W_hh: w[i] = W_hh[dir][kc*96 + i][j]
  float w[BCH];
  {
    const float* wbase = a.w_hh + (size_t)dir * G3 * BH + j;
#pragma unroll
    for (int i = 0; i < BCH; ++i) w[i] = __ldg(wbase + (size_t)(kc * BCH + i) * BH);
  }
  float* remote[BCL];
#pragma unroll
  for (int c = 0; c < BCL; ++c) remote[c] = cluster.map_shared_rank(gbuf, c);
  // where this lane's three d(gh) values live inside a row of the buffer
  int slot_g[3];
#pragma unroll
  for (int g = 0; g < 3; ++g) {
    const int gr = g * BH + j;
    slot_g[g] = (gr / BCH) * BCHP + (gr % BCH);
  }
  __syncthreads();
  cluster.sync();

  const int maxlen = lens[0];
  const int g_ld = a.dirs * G3, y_ld = a.dirs * BH;
  float dh_carry[BBT / 8];
#pragma unroll
  for (int i = 0; i < BBT / 8; ++i) dh_carry[i] = 0.f;

  for (int s = maxlen - 1; s >= 0; --s) {
    int cnt = (lens[lane] > s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    const int nact = cnt;
    const int ngroups = (nact + 7) >> 3;
    float dhz[BBT / 8];
    // ---- phase A: gate gradients of this CTA's units
#pragma unroll
    for (int rg = 0; rg < BBT / 8; ++rg) {
      dhz[rg] = 0.f;
      if (rg < ngroups) {
        const int row = rg * 8 + kc;
        if (row < nact) {
          const int len = lens[row];
          const int pos = (dir == 0) ? s : len - 1 - s;
          const int tok = toff[row] + pos;
          float dh = dh_carry[rg];
          if (a.dy) dh += a.dy[(size_t)tok * y_ld + dir * BH + j];
          if (s == len - 1 && a.dh_last) dh += a.dh_last[(size_t)rowid[row] * y_ld + dir * BH + j];
          const float* sv = a.saved + ((size_t)tok * a.dirs + dir) * 4 * BH + j;
          const float r = sv[0], z = sv[BH], n = sv[2 * BH], ghn = sv[3 * BH];
          float hprev = 0.f;
          if (s > 0) {
            const int tokp = toff[row] + ((dir == 0) ? s - 1 : len - s);
            hprev = a.y[(size_t)tokp * y_ld + dir * BH + j];
          }
          const float dn_pre = dh * (1.f - z) * (1.f - n * n);
          const float dz_pre = dh * (hprev - n) * z * (1.f - z);
          const float dr_pre = dn_pre * ghn * r * (1.f - r);
          const float dghn = dn_pre * r;
          float* gi = a.dgi + (size_t)tok * g_ld + dir * G3 + j;
          gi[0] = dr_pre; gi[BH] = dz_pre; gi[2 * BH] = dn_pre;
          float* gh = a.dgh + (size_t)tok * g_ld + dir * G3 + j;
          gh[0] = dr_pre; gh[BH] = dz_pre; gh[2 * BH] = dghn;
          const int base = row * BROW;
#pragma unroll
          for (int c = 0; c < BCL; ++c) {
            remote[c][base + slot_g[0]] = dr_pre;
            remote[c][base + slot_g[1]] = dz_pre;
            remote[c][base + slot_g[2]] = dghn;
          }
          dhz[rg] = dh * z;
        }
      }
    }
    cluster.sync();
    // ---- phase B: dh_{s-1}[own units] = dh*z + d(gh) . W_hh[:, own units]
#pragma unroll
    for (int rg = 0; rg < BBT / 8; ++rg) {
      if (rg < ngroups) {
        float acc[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) acc[r] = 0.f;
        const float* gb = gbuf + (rg * 8) * BROW + kc * BCHP;
#pragma unroll
        for (int i4 = 0; i4 < BCH / 4; ++i4) {
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            const float4 v = *reinterpret_cast<const float4*>(gb + r * BROW + i4 * 4);
            acc[r] = fmaf(v.x, w[i4 * 4 + 0], acc[r]);
            acc[r] = fmaf(v.y, w[i4 * 4 + 1], acc[r]);
            acc[r] = fmaf(v.z, w[i4 * 4 + 2], acc[r]);
            acc[r] = fmaf(v.w, w[i4 * 4 + 3], acc[r]);
          }
        }
        float a4[4], a2[2];
        {
          const bool up = (lane & 4) != 0;
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            float send = up ? acc[r] : acc[r + 4];
            float keep = up ? acc[r + 4] : acc[r];
            a4[r] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
          }
        }
        {
          const bool up = (lane & 2) != 0;
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            float send = up ? a4[r] : a4[r + 2];
            float keep = up ? a4[r + 2] : a4[r];
            a2[r] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
          }
        }
        const bool up = (lane & 1) != 0;
        const float send = up ? a2[0] : a2[1];
        const float keep = up ? a2[1] : a2[0];
        const float tot = keep + __shfl_xor_sync(0xffffffffu, send, 1);
        dh_carry[rg] = dhz[rg] + ((rg * 8 + kc) < nact ? tot : 0.f);
      }
    }
    cluster.sync();
  }
}

// ---- generic kernel: any H, 8 rows per CTA -------------------------------------------------
constexpr int BGEN_ROWS = 8;

__global__ void __launch_bounds__(256) gru_bwd_generic_kernel(GruBwdArgs a, int H) {
  extern __shared__ __align__(16) float smem[];
  float* dh = smem;                          // [BGEN_ROWS][H]  grad wrt h of the current step
  float* dhn = dh + BGEN_ROWS * H;           // [BGEN_ROWS][H]  next carry
  float* gbuf = dhn + BGEN_ROWS * H;         // [BGEN_ROWS][3H]
  __shared__ int lens[BGEN_ROWS], toff[BGEN_ROWS], rowid[BGEN_ROWS];
  const int dir = blockIdx.y;
  const int s0 = blockIdx.x * BGEN_ROWS;
  const int G3 = 3 * H;
  if (threadIdx.x < BGEN_ROWS) {
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
  for (int i = threadIdx.x; i < BGEN_ROWS * H; i += blockDim.x) dh[i] = 0.f;
  __syncthreads();
  const int maxlen = lens[0];
  const float* W = a.w_hh + (size_t)dir * G3 * H;
  const int g_ld = a.dirs * G3, y_ld = a.dirs * H;
  for (int s = maxlen - 1; s >= 0; --s) {
    for (int o = threadIdx.x; o < BGEN_ROWS * H; o += blockDim.x) {
      const int row = o / H, j = o % H;
      const int len = lens[row];
      float dr_pre = 0.f, dz_pre = 0.f, dghn = 0.f, dhzv = 0.f;
      if (s < len) {
        const int pos = (dir == 0) ? s : len - 1 - s;
        const int tok = toff[row] + pos;
        float d = dh[o];
        if (a.dy) d += a.dy[(size_t)tok * y_ld + dir * H + j];
        if (s == len - 1 && a.dh_last) d += a.dh_last[(size_t)rowid[row] * y_ld + dir * H + j];
        const float* sv = a.saved + ((size_t)tok * a.dirs + dir) * 4 * H + j;
        const float r = sv[0], z = sv[H], n = sv[2 * H], ghn = sv[3 * H];
        float hprev = 0.f;
        if (s > 0) {
          const int tokp = toff[row] + ((dir == 0) ? s - 1 : len - s);
          hprev = a.y[(size_t)tokp * y_ld + dir * H + j];
        }
        const float dn_pre = d * (1.f - z) * (1.f - n * n);
        dz_pre = d * (hprev - n) * z * (1.f - z);
        dr_pre = dn_pre * ghn * r * (1.f - r);
        dghn = dn_pre * r;
        float* gi = a.dgi + (size_t)tok * g_ld + dir * G3 + j;
        gi[0] = dr_pre; gi[H] = dz_pre; gi[2 * H] = dn_pre;
        float* gh = a.dgh + (size_t)tok * g_ld + dir * G3 + j;
        gh[0] = dr_pre; gh[H] = dz_pre; gh[2 * H] = dghn;
        dhzv = d * z;
      }
      gbuf[row * G3 + j] = dr_pre;
      gbuf[row * G3 + H + j] = dz_pre;
      gbuf[row * G3 + 2 * H + j] = dghn;
      dhn[o] = dhzv;
    }
    __syncthreads();
    for (int o = threadIdx.x; o < BGEN_ROWS * H; o += blockDim.x) {
      const int row = o / H, k = o % H;
      float acc = dhn[o];
      if (s < lens[row]) {
        const float* gb = gbuf + row * G3;
        for (int gr = 0; gr < G3; ++gr) acc = fmaf(gb[gr], __ldg(W + (size_t)gr * H + k), acc);
      }
      dh[o] = acc;
    }
    __syncthreads();
  }
}

// hprev[tok(step s)] = y[tok(step s-1)] (0 at the first step) for every direction.  One CTA per row, one
// 16-byte column chunk per thread (no per-element div/mod: the first version spent 0.33 ms per call on 0.57 GB).
__global__ void __launch_bounds__(128)
gru_hprev_kernel(const float* __restrict__ y, const int32_t* __restrict__ offsets, int H, int dirs,
                 float* __restrict__ hprev) {
  const int s = blockIdx.x;
  const int off = offsets[s], len = offsets[s + 1] - off;
  const int ld = dirs * H;
  if ((H & 3) == 0) {
    const int ld4 = ld >> 2;
    const float4* y4 = reinterpret_cast<const float4*>(y);
    float4* h4 = reinterpret_cast<float4*>(hprev);
    for (int c = threadIdx.x; c < ld4; c += blockDim.x) {
      const int step = ((c << 2) / H == 0) ? -1 : 1;            // forward direction looks one token back, reverse one ahead
      for (int t = 0; t < len; ++t) {
        const int tp = t + step;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (tp >= 0 && tp < len) v = __ldg(y4 + (size_t)(off + tp) * ld4 + c);
        h4[(size_t)(off + t) * ld4 + c] = v;
      }
    }
    return;
  }
  for (int i = threadIdx.x; i < len * ld; i += blockDim.x) {
    const int t = i / ld, c = i % ld;
    const int dir = c / H;
    const int tp = (dir == 0) ? t - 1 : t + 1;
    float v = 0.f;
    if (tp >= 0 && tp < len) v = y[(size_t)(off + tp) * ld + c];
    hprev[(size_t)(off + t) * ld + c] = v;
  }
}

int64_t gru_bwd_tc_workspace_bytes(int B, int dirs);
int launch_gru_bwd_tc(const float* dy, const float* dh_last, const float* y, const float* saved, const float* w_hh,
                      const int32_t* order, const int32_t* offsets, int B, int dirs, float* dgi, float* dgh,
                      void* workspace, float* db_ih, float* db_hh, cudaStream_t st);

}  // namespace ttr

extern "C" int64_t ttr_gru_bwd_workspace_bytes(int B, int H, int dirs) {
  return H == ttr::BH ? ttr::gru_bwd_tc_workspace_bytes(B, dirs) : 0;
}

extern "C" int ttr_gru_recurrence_bwd(const float* dy, const float* dh_last, const float* y, const float* saved,
                                      const float* w_hh, const int32_t* order, const int32_t* offsets, int B, int H,
                                      int dirs, float* dgi, float* dgh, void* stream);

extern "C" int ttr_gru_recurrence_bwd_ws(const float* dy, const float* dh_last, const float* y, const float* saved,
                                         const float* w_hh, const int32_t* order, const int32_t* offsets, int B, int H,
                                         int dirs, float* dgi, float* dgh, void* workspace, int64_t workspace_bytes,
                                         int m_bound, const int32_t* m_valid, float* db_ih, float* db_hh,
                                         void* stream) {
  using namespace ttr;
  TTR_REQUIRE(B >= 1 && H >= 1 && (dirs == 1 || dirs == 2), "ttr_gru_recurrence_bwd_ws: bad shape");
  TTR_REQUIRE(y && saved && dgi && dgh, "ttr_gru_recurrence_bwd_ws: y, saved, dgi, dgh are required");
  if (H == BH && workspace != nullptr && !(g_debug_flags & (1 | (1 << 23)))) {
    TTR_REQUIRE(workspace_bytes >= gru_bwd_tc_workspace_bytes(B, dirs), "ttr_gru_recurrence_bwd_ws: workspace of %lld B < %lld B",
                (long long)workspace_bytes, (long long)gru_bwd_tc_workspace_bytes(B, dirs));
    return launch_gru_bwd_tc(dy, dh_last, y, saved, w_hh, order, offsets, B, dirs, dgi, dgh, workspace, db_ih, db_hh,
                             (cudaStream_t)stream);
  }
  int rc = ttr_gru_recurrence_bwd(dy, dh_last, y, saved, w_hh, order, offsets, B, H, dirs, dgi, dgh, stream);
  // the other kernels leave the bias gradients to the column-sum kernel (accumulating, like the fused path)
  if (rc == TTR_OK && db_hh) rc = ttr_colsum(dgh, m_bound, m_valid, dirs * 3 * H, db_hh, 1, stream);
  if (rc == TTR_OK && db_ih) rc = ttr_colsum(dgi, m_bound, m_valid, dirs * 3 * H, db_ih, 1, stream);
  return rc;
}

extern "C" int ttr_gru_recurrence_bwd(const float* dy, const float* dh_last, const float* y, const float* saved,
                                      const float* w_hh, const int32_t* order, const int32_t* offsets, int B, int H,
                                      int dirs, float* dgi, float* dgh, void* stream) {
  using namespace ttr;
  TTR_REQUIRE(B >= 1 && H >= 1 && (dirs == 1 || dirs == 2), "ttr_gru_recurrence_bwd: bad shape");
  TTR_REQUIRE(y && saved && dgi && dgh, "ttr_gru_recurrence_bwd: y, saved, dgi, dgh are required");
  cudaStream_t st = (cudaStream_t)stream;
  GruBwdArgs a{dy, dh_last, y, saved, w_hh, order, offsets, B, dirs, dgi, dgh};
  if (H == BH && !(g_debug_flags & 1)) {
    const size_t smem = (size_t)BBT * BROW * sizeof(float) + 3 * BBT * sizeof(int);
    TTR_CHECK_CUDA(cudaFuncSetAttribute(gru_bwd_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(ceil_div(B, BBT) * BCL, dirs);
    gru_bwd_cluster_kernel<<<grid, BTHREADS, smem, st>>>(a);
    TTR_CHECK_LAUNCH();
  } else {
    const size_t smem = (size_t)BGEN_ROWS * 5 * H * sizeof(float);
    TTR_REQUIRE(smem <= 200 * 1024, "ttr_gru_recurrence_bwd: H=%d too large for the generic kernel", H);
    if (smem > 48 * 1024)
      TTR_CHECK_CUDA(cudaFuncSetAttribute(gru_bwd_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(ceil_div(B, BGEN_ROWS), dirs);
    gru_bwd_generic_kernel<<<grid, 256, smem, st>>>(a, H);
    TTR_CHECK_LAUNCH();
  }
  return TTR_OK;
}

extern "C" int ttr_gru_whh_grad(const float* dgh, const float* y, const int32_t* offsets, int B, int H, int dirs,
                                int m_bound, float* hprev_ws, float* dw_hh, int accumulate, void* stream) {
  using namespace ttr;
  TTR_REQUIRE(B >= 1 && H >= 1 && (dirs == 1 || dirs == 2) && m_bound >= 1, "ttr_gru_whh_grad: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  gru_hprev_kernel<<<B, 128, 0, st>>>(y, offsets, H, dirs, hprev_ws);
  TTR_CHECK_LAUNCH();
  const int G3 = 3 * H;
  int rc = ttr_zero_tail_rows(hprev_ws, m_bound, offsets + B, dirs * H, stream);
  if (rc != TTR_OK) return rc;
  for (int dir = 0; dir < dirs; ++dir) {
    // dW_hh[dir] [3H, H] (+)= dgh[:, dir*3H : (dir+1)*3H]^T * hprev[:, dir*H : (dir+1)*H]   (tcgen05, split-K)
    rc = ttr_gemm_tn_tf32(dgh + dir * G3, dirs * G3, hprev_ws + dir * H, dirs * H, dw_hh + (size_t)dir * G3 * H, H,
                          m_bound, offsets + B, G3, H, accumulate, stream);
    if (rc != TTR_OK) return rc;
  }
  return TTR_OK;
}
