// gru_fwd_tc.cu — GRU recurrence (forward, H = 256) with the recurrent product on tcgen05.
//
// Replaces the sequential half of `self.rnn(packed)` (backend/model.py:59-62), like gru_fwd.cu,
// but moves gh_t = h_{t-1} W_hh^T from fp32 FMAs + warp shuffles to the tensor cores:
//
//   * a thread-block cluster of 8 CTAs owns one tile of 256 length-sorted rows and one direction
//     for ALL timesteps, as two independent 128-row chains (UMMA M = 128) that interleave on the
//     SM: while one chain's h is in flight between CTAs the other one computes.
//     CTA c owns hidden units [32c, 32c+32), i.e. 96 gate columns (r, z, n).
//   * its slice of W_hh, rounded to fp16, lives in shared memory for the whole kernel as the UMMA
//     B operand (48 KB); h_{t-1} of each chain, rounded to fp16, is the A operand (64 KB per chain).
//     One kind::f16 MMA chain of 16 k-steps per timestep accumulates in fp32 in tensor memory.
//     Precision: the fp32 state h is carried in registers by the thread that owns it; only the
//     matmul INPUTS are fp16 (11-bit significand, finer than the tf32 operands of the input
//     projection).  Simulated on the reference model at config dims the rounding moves the final
//     embeddings by 7.9e-5 relative on its own and 3.30e-4 -> 3.33e-4 together with the tf32
//     projection (tolerance 1e-3); a three-plane hi/lo split version of this kernel (fp32-exact to
//     2^-22) measured the same parity and 3x the tensor and exchange cost — see git history.
//   * epilogue thread = (row, 16 units): tcgen05.ld of its 48 gate pre-activations, fused gate
//     math against the gi row segment it prefetched while the MMAs ran, results written to
//     y / saved / h_last, and the new h values written as fp16 straight into the CTA's own slice
//     of the A operand.
//   * exchange: one elected thread copies the CTA's 8 KB slice from its own A operand to a
//     double-buffered global scratch image of the operand (L2-resident) with a bulk store, waits for
//     that store, and issues ONE cp.async.bulk ... multicast::cluster load that lands the slice in
//     the seven peers' A operands and completes on each destination's mbarrier — no cluster-wide
//     barrier in the loop and 8 KB instead of 56 KB leaving the SM per step.  (Writing the scratch
//     with plain stores needed a gpu-scope fence in every epilogue thread: 2.3 k cycles per step.)  (The first version pushed the
//     slice with shared::cta -> shared::cluster bulk copies: 15-30 B/clk per SM, the longest phase
//     of the step, profiles/r1_gru_tc_trace_v1.txt.)  The write-after-read hazard (a peer's MMA
//     still reading its operand) is covered by a second mbarrier that every CTA's MMA completion
//     signals in all eight CTAs (tcgen05.commit ... multicast::cluster); the same signal, one step
//     later, proves that a scratch slice has been delivered everywhere before it is rewritten.
//
// Operand layout (both operands K-major, no swizzle): [k / 8][row][8 halves] — 8-row x 16-byte core
// matrices, stride-byte-offset 128 B between row groups, leading-byte-offset rows*16 B between
// k-chunks — so a thread's 16-byte store lands conflict-free and a CTA's unit slice (4 k-chunks)
// is one contiguous 8 KB block for the bulk copy.
#include <cuda_fp16.h>

#include "common.cuh"
#include "ptx.cuh"

namespace ttr {

extern int g_debug_flags;
extern long long* g_score_trace;     // ttr_debug_set_trace: also receives this kernel's per-step timeline

constexpr int TC_H = 256;
constexpr int TC_CL = 8;                         // CTAs per cluster
constexpr int TC_UN = TC_H / TC_CL;              // 32 hidden units per CTA
constexpr int TC_NG = 3 * TC_UN;                 // 96 gate columns per CTA (UMMA N)
constexpr int TC_ROWS = 128;                     // rows per chain (UMMA M)
constexpr int TC_CHAINS = 2;                     // independent chains per cluster
constexpr int TC_TILE = TC_ROWS * TC_CHAINS;     // rows per cluster
constexpr int TC_KC = TC_H / 8;                  // 32 k-chunks of 8 halves
constexpr int TC_A_LBO = TC_ROWS * 16;           // 2048 B between k-chunks of the A operand
constexpr int TC_B_LBO = TC_NG * 16;             // 1536 B between k-chunks of the B operand
constexpr int TC_A_BYTES = TC_KC * TC_A_LBO;     // 64 KB per chain
constexpr int TC_B_BYTES = TC_KC * TC_B_LBO;     // 48 KB
constexpr int TC_SLICE_BYTES = (TC_UN / 8) * TC_A_LBO;   // 8 KB: one CTA's units of one chain
constexpr int TC_GROUP = 256;                    // threads per chain
constexpr int TC_THREADS = TC_GROUP * TC_CHAINS;
constexpr int TC_ACC_COLS = 128;                 // TMEM columns per chain (96 used)
constexpr int TC_TMEM_COLS = TC_ACC_COLS * TC_CHAINS;
constexpr int TC_SMEM = TC_CHAINS * TC_A_BYTES + TC_B_BYTES + TC_NG * 4 + 3 * TC_TILE * 4 + TC_CHAINS * 3 * 8 + 8;

struct GruTcArgs {
  const float* gi;          // HALF_IO: const __half*
  const float* w_hh;
  const float* b_hh;
  const int32_t* order;
  const int32_t* offsets;
  int B, dirs;
  float* y;                 // HALF_IO: __half*
  float* h_last;
  float* saved;
  unsigned char* scratch;   // [clusters][chains][2][TC_A_BYTES] operand images for the multicast exchange
  int tile_rows;            // rows per cluster: TC_TILE (two chains) or TC_ROWS (one chain; small batches)
  long long* trace;   // diagnostic: [8][256] clock64 of cluster 0 / CTA 0 at eight points of each step, or NULL
};

#define TC_TRACE(role)                                                                     \
  do {                                                                                     \
    if (a.trace && blockIdx.x == 0 && blockIdx.y == 0 && t < 256) a.trace[(role) * 256 + t] = clock64(); \
  } while (0)

namespace {

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// global -> shared bulk copy delivered to the same CTA-relative address (and mbarrier) in every CTA of `mask`
__device__ __forceinline__ void bulk_g2s_multicast(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar,
                                                   uint16_t mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::
          "r"(dst),
      "l"(src), "r"(bytes), "r"(bar), "h"(mask)
      : "memory");
}
// shared::cta -> global bulk copy (bulk async-group completion)
__device__ __forceinline__ void bulk_s2g(void* gdst, uint32_t ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(ssrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mma_commit_multicast(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          ptx::smem_u32(bar)),
      "h"(mask)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
      : "r"(taddr)
      : "memory");
}
// K-major operand without swizzle: 8-row x 16-byte core matrices; LBO = distance between the two
// k-chunks of one MMA (and of consecutive k-chunks), SBO = distance between 8-row groups.
__device__ __forceinline__ uint64_t make_kmajor_nosw_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;      // descriptor version (Blackwell); layout type 0 = no swizzle
  return d;
}
// 8 fp32 values -> 8 fp16 (round to nearest even) in one 16-byte row of an operand
__device__ __forceinline__ uint4 pack8(const float* x) {
  uint32_t h[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __half2 v = __floats2half2_rn(x[2 * i], x[2 * i + 1]);
    h[i] = *reinterpret_cast<const uint32_t*>(&v);
  }
  return make_uint4(h[0], h[1], h[2], h[3]);
}

// 4x4 transpose of 16-byte elements inside each group of 4 lanes: on entry lane c of a group holds
// v[k] = M[k][c]; on exit it holds v[k] = M[c][k].  Used to turn "lane = row" register tiles into
// row-contiguous global accesses: a warp-wide load where every lane reads 16 B of its own row touches 32
// lines (32 L1 wavefronts); with lane (G, c) reading chunk c of row 4G + k it touches 8 rows x 64 B.
__device__ __forceinline__ float4 shfl_xor_f4(const float4& v, int m) {
  return make_float4(__shfl_xor_sync(0xffffffffu, v.x, m), __shfl_xor_sync(0xffffffffu, v.y, m),
                     __shfl_xor_sync(0xffffffffu, v.z, m), __shfl_xor_sync(0xffffffffu, v.w, m));
}
__device__ __forceinline__ void transpose4(float4 (&v)[4], int lane) {
  {
    const bool hi = (lane & 1) != 0;
    const float4 r0 = shfl_xor_f4(hi ? v[0] : v[1], 1);
    const float4 r1 = shfl_xor_f4(hi ? v[2] : v[3], 1);
    if (hi) { v[0] = r0; v[2] = r1; } else { v[1] = r0; v[3] = r1; }
  }
  {
    const bool hi = (lane & 2) != 0;
    const float4 r0 = shfl_xor_f4(hi ? v[0] : v[2], 2);
    const float4 r1 = shfl_xor_f4(hi ? v[1] : v[3], 2);
    if (hi) { v[0] = r0; v[1] = r1; } else { v[2] = r0; v[3] = r1; }
  }
}

// Activations on the MUFU pipe (16 lanes/clk/SM: the epilogue's scarcest unit) share reciprocals:
// sigmoid(a), sigmoid(b) = (1+e^-b, 1+e^-a) / ((1+e^-a)(1+e^-b)) — 2 ex2 + 1 rcp; two tanh likewise.
// Arguments are clamped from below (-20 / -10) so that products of denominators stay finite
// (sigmoid(-20) = 2e-9, tanh(-10) = -1 + 4e-9); large positive arguments just make e^-x vanish.
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void sigmoid2(float a, float b, float& sa, float& sb) {
  const float da = 1.0f + __expf(-fmaxf(a, -20.0f)), db = 1.0f + __expf(-fmaxf(b, -20.0f));
  const float inv = rcp_approx(da * db);
  sa = db * inv;
  sb = da * inv;
}
__device__ __forceinline__ void tanh2(float a, float b, float& ta, float& tb) {
  const float ea = __expf(-2.0f * fmaxf(a, -10.0f)), eb = __expf(-2.0f * fmaxf(b, -10.0f));   // tanh(x) = (1 - e^-2x) / (1 + e^-2x)
  const float da = 1.0f + ea, db = 1.0f + eb;
  const float inv = rcp_approx(da * db);
  ta = (1.0f - ea) * db * inv;
  tb = (1.0f - eb) * da * inv;
}

// per-step outputs y[tok, col0 + 16 units] of a warp's 32 rows, written row-contiguous (see transpose4);
// HALF: y holds fp16 (it only feeds the next layer's kind::f16 projection)
template <bool HALF>
__device__ __forceinline__ void store_y(float* y, const float (&h)[16], const int (&tok_k)[4], int t,
                                        const int* lens4, int y_ld, int col0, int lane) {
  if (y == nullptr) return;
  float4 v[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) v[k] = make_float4(h[4 * k], h[4 * k + 1], h[4 * k + 2], h[4 * k + 3]);
  transpose4(v, lane);                         // lane (G, c) now holds units 4c..4c+3 of rows 4G + k
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (t >= lens4[k]) continue;
    if (HALF) {
      const __half2 lo = __floats2half2_rn(v[k].x, v[k].y), hi = __floats2half2_rn(v[k].z, v[k].w);
      *(reinterpret_cast<uint2*>(reinterpret_cast<__half*>(y) + (size_t)tok_k[k] * y_ld + col0) + (lane & 3)) =
          make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
    } else {
      *(reinterpret_cast<float4*>(y + (size_t)tok_k[k] * y_ld + col0) + (lane & 3)) = v[k];
    }
  }
}

}  // namespace

// HALF_IO: gi is read and y written as fp16 (the inference pipeline; `saved` must then be NULL)
template <bool HALF_IO>
__global__ void __cluster_dims__(TC_CL, 1, 1) __launch_bounds__(TC_THREADS, 1)
gru_fwd_tc_kernel(GruTcArgs a) {
  extern __shared__ __align__(128) unsigned char sm[];
  unsigned char* h_all = sm;                                        // [chains] A operands
  unsigned char* w_sm = sm + TC_CHAINS * TC_A_BYTES;                // B operand
  float* bias = reinterpret_cast<float*>(w_sm + TC_B_BYTES);        // [96]: b_hr, b_hz, b_hn of the slice
  int* lens = reinterpret_cast<int*>(bias + TC_NG);
  int* toff = lens + TC_TILE;
  int* rowid = toff + TC_TILE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(rowid + TC_TILE);    // per chain: h_full, mma_done, consumed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + TC_CHAINS * 3);

  // warp index through a shuffle: ptxas then knows it is warp-uniform and keeps everything derived from it
  // (chain, operand addresses, TMEM columns) in uniform registers — otherwise every tcgen05.mma operand
  // goes through an R2UR waterfall loop (~100 cycles per MMA instead of ~60)
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = tid & 31;
  const int rank = (int)cluster_ctarank();
  const int tile = blockIdx.x / TC_CL;
  const int dir = blockIdx.y;
  const int G3 = 3 * TC_H;
  const int s0 = tile * a.tile_rows;

  if (tid == 0) {
    for (int c = 0; c < TC_CHAINS; ++c) {
      ptx::mbar_init(bars + 3 * c + 0, 1);        // h_full: peers' slices have landed
      ptx::mbar_init(bars + 3 * c + 1, 1);        // mma_done: own accumulators are ready
      ptx::mbar_init(bars + 3 * c + 2, TC_CL);    // consumed: all 8 CTAs finished reading h_{t-1}
    }
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, TC_TMEM_COLS);
    ptx::tmem_relinquish();
  }
  for (int i = tid; i < TC_TILE; i += TC_THREADS) {
    const int s = s0 + i;
    if (s < a.B && i < a.tile_rows) {
      const int off = a.offsets[s];
      lens[i] = a.offsets[s + 1] - off;
      toff[i] = off;
      rowid[i] = a.order[s];
    } else {
      lens[i] = 0; toff[i] = 0; rowid[i] = 0;
    }
  }
  {  // h_0 = 0
    uint4* p = reinterpret_cast<uint4*>(h_all);
    for (int i = tid; i < TC_CHAINS * TC_A_BYTES / 16; i += TC_THREADS) p[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  {  // resident weights: rows (gate g, unit 32*rank + u) of W_hh[dir] as fp16
    const float* wbase = a.w_hh + (size_t)dir * G3 * TC_H;
    for (int idx = tid; idx < TC_NG * TC_KC; idx += TC_THREADS) {
      const int n = idx % TC_NG, kc = idx / TC_NG;
      const int g = n / TC_UN, u = n % TC_UN;
      const float4* src = reinterpret_cast<const float4*>(wbase + (size_t)(g * TC_H + rank * TC_UN + u) * TC_H + kc * 8);
      const float4 v0 = __ldg(src), v1 = __ldg(src + 1);
      const float x[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
      *reinterpret_cast<uint4*>(w_sm + kc * TC_B_LBO + n * 16) = pack8(x);
    }
    if (tid < TC_NG) bias[tid] = a.b_hh[dir * G3 + (tid / TC_UN) * TC_H + rank * TC_UN + (tid % TC_UN)];
  }
  ptx::fence_proxy_async_smem();          // operand bytes written by threads -> visible to the tensor core
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  cluster_sync_all();                     // every CTA's barriers are initialised before any remote arrive
  const uint32_t tmem_base = *tmem_slot;

  // role of this thread: chain `ch`; inside the chain TMEM lane quarter q (rows 32q..32q+31), unit half uh
  const int ch = warp >> 3, wg = warp & 7;
  const int q = wg & 3, uh = wg >> 2;
  const int row = q * 32 + lane;                    // row inside the chain
  const int u0 = uh * 16;                           // first owned unit inside the CTA's slice
  const int j0 = rank * TC_UN + u0;                 // ... as a global hidden unit
  unsigned char* h_sm = h_all + ch * TC_A_BYTES;
  uint64_t* h_full = bars + 3 * ch;
  uint64_t* mma_done = h_full + 1;
  uint64_t* consumed = h_full + 2;
  const uint32_t tmem_acc = tmem_base + (uint32_t)(ch * TC_ACC_COLS);
  const uint32_t tmem_row = tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)u0;
  const int len = lens[ch * TC_ROWS + row];
  const int tbase = toff[ch * TC_ROWS + row];
  const int maxlen = lens[ch * TC_ROWS];
  const int gi_ld = a.dirs * G3, y_ld = a.dirs * TC_H;
  const float* gi_base = a.gi + dir * G3 + j0;                                   // fp32 view
  const __half* gi16_base = reinterpret_cast<const __half*>(a.gi) + dir * G3 + j0;   // fp16 view
  float h[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) h[i] = 0.f;

  const uint32_t h_u32 = ptx::smem_u32(h_sm);
  unsigned char* scr = a.scratch + ((size_t)(blockIdx.y * (gridDim.x / TC_CL) + tile) * TC_CHAINS + ch) * 2 * TC_A_BYTES;
  const uint32_t w_u32 = ptx::smem_u32(w_sm);
  constexpr uint32_t idesc = ptx::make_idesc_f16(TC_ROWS, TC_NG);
  const int kc0 = 4 * rank + 2 * uh;                // k-chunk of this thread's first 8 units

  for (int t = 0; t < maxlen; ++t) {
    const uint32_t par = (uint32_t)t & 1u;
    const bool active = t < len;
    const int tok = active ? tbase + (dir == 0 ? t : len - 1 - t) : 0;

    if (wg == 0 && lane == 0 && ch == 0) TC_TRACE(0);
    if (wg == 0) {
      // ===== MMA issue: gh = h W_hh^T for this chain (16 k-steps of 16) =====
      if (t > 0) ptx::mbar_wait(h_full, par ^ 1u);
      ptx::tc_fence_after_sync();
      if (lane == 0 && ch == 0) TC_TRACE(1);
      if (ptx::elect_one()) {
        const uint64_t a_desc = make_kmajor_nosw_desc(h_u32, TC_A_LBO, 128);
        const uint64_t b_desc = make_kmajor_nosw_desc(w_u32, TC_B_LBO, 128);
#pragma unroll
        for (int ks = 0; ks < TC_H / 16; ++ks)
          ptx::mma_f16_ss(tmem_acc, a_desc + (uint64_t)(ks * (2 * TC_A_LBO >> 4)), b_desc + (uint64_t)(ks * (2 * TC_B_LBO >> 4)),
                     idesc, ks != 0);
        ptx::mma_commit(mma_done);
        // nobody waits for the last step's signal, and a peer may have left by the time it would land
        if (t + 1 < maxlen) mma_commit_multicast(consumed, (uint16_t)0xff);
      }
      __syncwarp();
      if (lane == 0 && ch == 0) TC_TRACE(2);
    }

    // gi of this step (3 gates x 16 units per thread), in flight while the MMAs run.  Loaded row-contiguous:
    // lane (G, c) reads units 4c..4c+3 of rows 4G + k (k = 0..3), transposed into place after the wait.
    // Rows past their end read token 0 (valid memory, unused).
    const int grp4 = lane & ~3, c4 = lane & 3;
    int tok_k[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) tok_k[k] = __shfl_sync(0xffffffffu, tok, grp4 + k);
    float4 g4[3][4];
#pragma unroll
    for (int g = 0; g < 3; ++g)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (HALF_IO) {
          const uint2 raw = __ldg(reinterpret_cast<const uint2*>(gi16_base + (size_t)tok_k[k] * gi_ld + g * TC_H) + c4);
          const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&raw.x));
          const float2 hi = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
          g4[g][k] = make_float4(lo.x, lo.y, hi.x, hi.y);
        } else {
          g4[g][k] = __ldg(reinterpret_cast<const float4*>(gi_base + (size_t)tok_k[k] * gi_ld + g * TC_H) + c4);
        }
      }

    ptx::mbar_wait(mma_done, par);
    ptx::tc_fence_after_sync();
    if (wg == 0 && lane == 0 && ch == 0) TC_TRACE(3);
#pragma unroll
    for (int g = 0; g < 3; ++g) transpose4(g4[g], lane);
    if (__any_sync(0xffffffffu, active)) {
      const float* gf = reinterpret_cast<const float*>(g4);   // [gate][16 units] of this thread's row
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        uint32_t ar[8], az[8], an[8];
        tmem_ld_32x8(tmem_row + 8 * hf, ar);
        tmem_ld_32x8(tmem_row + TC_UN + 8 * hf, az);
        tmem_ld_32x8(tmem_row + 2 * TC_UN + 8 * hf, an);
        ptx::tmem_ld_wait();
        if (active) {
          float rr[8], zz[8], nn[8], gn[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int u = 8 * hf + i;
            sigmoid2(gf[u] + __uint_as_float(ar[i]) + bias[u0 + u],
                     gf[16 + u] + __uint_as_float(az[i]) + bias[TC_UN + u0 + u], rr[i], zz[i]);
            gn[i] = __uint_as_float(an[i]) + bias[2 * TC_UN + u0 + u];
          }
#pragma unroll
          for (int i = 0; i < 8; i += 2) {
            tanh2(fmaf(rr[i], gn[i], gf[32 + 8 * hf + i]), fmaf(rr[i + 1], gn[i + 1], gf[32 + 8 * hf + i + 1]), nn[i],
                  nn[i + 1]);
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int u = 8 * hf + i;
            h[u] = fmaf(zz[i], h[u] - nn[i], nn[i]);            // (1-z)*n + z*h
          }
          // new h (fp16) -> own slice of the A operand (k = 32*rank + u0 + 8*hf + i); the issuing thread copies the
          // slice to the scratch image and multicasts it to the peers
          *reinterpret_cast<uint4*>(h_sm + (kc0 + hf) * TC_A_LBO + row * 16) = pack8(h + 8 * hf);
          if (a.saved) {
            float4* sv = reinterpret_cast<float4*>(a.saved + ((size_t)tok * a.dirs + dir) * 4 * TC_H + j0 + 8 * hf);
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              sv[i] = make_float4(rr[4 * i], rr[4 * i + 1], rr[4 * i + 2], rr[4 * i + 3]);
              sv[TC_H / 4 + i] = make_float4(zz[4 * i], zz[4 * i + 1], zz[4 * i + 2], zz[4 * i + 3]);
              sv[2 * TC_H / 4 + i] = make_float4(nn[4 * i], nn[4 * i + 1], nn[4 * i + 2], nn[4 * i + 3]);
              sv[3 * TC_H / 4 + i] = make_float4(gn[4 * i], gn[4 * i + 1], gn[4 * i + 2], gn[4 * i + 3]);
            }
          }
        }
      }
    }
    if (wg == 0 && lane == 0 && ch == 0) TC_TRACE(4);
    ptx::tc_fence_before_sync();
    ptx::fence_proxy_async_smem();        // the slice written above is read by the copy engine and by the tensor core
    ptx::named_bar_sync(1 + ch, TC_GROUP);
    if (wg == 0 && lane == 0 && ch == 0) TC_TRACE(5);
    // the exchange is issued first (below, by one thread of warp 0); the per-step outputs leave after the
    // fence so that it only waits for the 32 bytes of scratch per thread, not for the HBM writes
    if (wg != 0) store_y<HALF_IO>(a.y, h, tok_k, t, lens + ch * TC_ROWS + q * 32 + grp4, y_ld, dir * TC_H + j0, lane);
    if (active && wg != 0) {
      if (t == len - 1) {
        float4* hp = reinterpret_cast<float4*>(a.h_last + (size_t)rowid[ch * TC_ROWS + row] * y_ld + dir * TC_H + j0);
#pragma unroll
        for (int i = 0; i < 4; ++i) hp[i] = make_float4(h[4 * i], h[4 * i + 1], h[4 * i + 2], h[4 * i + 3]);
      }
    }

    if (wg == 0 && t + 1 < maxlen) {
      // ===== exchange: once all 8 CTAs have finished reading h_{t-1}, land the new slice in all of them =====
      if (ptx::elect_one()) {
        // own slice: smem -> scratch image in L2 by the copy engine; its completion orders it before the multicast
        // load below (a per-thread gpu-scope fence after plain stores cost ~2.3 k cycles of the step here)
        const uint32_t off = (uint32_t)rank * TC_SLICE_BYTES;
        bulk_s2g(scr + (size_t)par * TC_A_BYTES + off, h_u32 + off, TC_SLICE_BYTES);
        ptx::bulk_commit_group();
        ptx::bulk_wait_group<0>();
        ptx::mbar_wait(consumed, par);
        if (ch == 0) TC_TRACE(6);
        ptx::mbar_arrive_expect_tx(h_full, (TC_CL - 1) * TC_SLICE_BYTES);      // seven slices arrive here, one per peer
        bulk_g2s_multicast(h_u32 + off, scr + (size_t)par * TC_A_BYTES + off, TC_SLICE_BYTES, ptx::smem_u32(h_full),
                           (uint16_t)(0xff & ~(1u << rank)));
        if (ch == 0) TC_TRACE(7);
      }
      __syncwarp();
    }
    if (wg == 0) store_y<HALF_IO>(a.y, h, tok_k, t, lens + ch * TC_ROWS + q * 32 + grp4, y_ld, dir * TC_H + j0, lane);
    if (active && wg == 0) {
      if (t == len - 1) {
        float4* hp = reinterpret_cast<float4*>(a.h_last + (size_t)rowid[ch * TC_ROWS + row] * y_ld + dir * TC_H + j0);
#pragma unroll
        for (int i = 0; i < 4; ++i) hp[i] = make_float4(h[4 * i], h[4 * i + 1], h[4 * i + 2], h[4 * i + 3]);
      }
    }
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();                     // no CTA leaves while a peer's commit may still arrive on its barriers
  if (warp == 1) ptx::tmem_dealloc(tmem_base, TC_TMEM_COLS);
}

// Steps are latency chains, so when every cluster of 128-row tiles is co-resident anyway (<= 15 of them) one
// chain per cluster on twice as many SMs beats two interleaved chains per cluster.
// More generally: the fewest rows per cluster (>= 128, <= 256, multiple of 8) that keeps the whole batch within
// the 15 clusters of 8 a B200 holds at once; chain 1 then carries the rows beyond 128 (possibly only a few).
static int tc_tile_rows(int B, int dirs) {
  const int slots = 15 / dirs;                                   // cluster tiles per direction in one wave
  const int need = ceil_div(ceil_div(B, slots), 8) * 8;
  return need <= TC_ROWS ? TC_ROWS : (need <= TC_TILE ? need : TC_TILE);
}

int64_t gru_fwd_tc_workspace_bytes(int B, int dirs) {
  return (int64_t)ceil_div(B, TC_ROWS) * dirs * TC_CHAINS * 2 * TC_A_BYTES;      // sized for either tiling
}

int launch_gru_fwd_tc(const float* gi, const float* w_hh, const float* b_hh, const int32_t* order,
                      const int32_t* offsets, int B, int dirs, float* y, float* h_last, float* saved, void* workspace,
                      bool half_io, cudaStream_t st) {
  const int tile_rows = tc_tile_rows(B, dirs);
  GruTcArgs a{gi, w_hh, b_hh, order, offsets, B, dirs, y, h_last, saved, reinterpret_cast<unsigned char*>(workspace),
              tile_rows, g_score_trace};
  dim3 grid(ceil_div(B, tile_rows) * TC_CL, dirs);
  if (half_io) {
    TTR_REQUIRE(saved == nullptr, "tcgen05 GRU with fp16 gi/y is inference-only (saved must be NULL)");
    TTR_CHECK_CUDA(cudaFuncSetAttribute(gru_fwd_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM));
    gru_fwd_tc_kernel<true><<<grid, TC_THREADS, TC_SMEM, st>>>(a);
    TTR_CHECK_LAUNCH();
    return TTR_OK;
  }
  TTR_CHECK_CUDA(cudaFuncSetAttribute(gru_fwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM));
  gru_fwd_tc_kernel<false><<<grid, TC_THREADS, TC_SMEM, st>>>(a);
  TTR_CHECK_LAUNCH();
  return TTR_OK;
}

}  // namespace ttr

// diagnostic: how many 8-CTA clusters of the tcgen05 recurrence the device can hold at once
extern "C" int ttr_debug_gru_tc_max_clusters(int* out) {
  using namespace ttr;
  TTR_CHECK_CUDA(cudaFuncSetAttribute(gru_fwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(TC_CL * 64, 1, 1);
  cfg.blockDim = dim3(TC_THREADS, 1, 1);
  cfg.dynamicSmemBytes = TC_SMEM;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeClusterDimension;
  attr.val.clusterDim.x = TC_CL; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
  cfg.attrs = &attr; cfg.numAttrs = 1;
  TTR_CHECK_CUDA(cudaOccupancyMaxActiveClusters(out, gru_fwd_tc_kernel<false>, &cfg));
  return TTR_OK;
}
