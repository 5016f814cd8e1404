// gru_fwd_tc.cu — GRU recurrence (forward, H = 256) with the recurrent product on tcgen05.
//
// Replaces the sequential half of `self.rnn(packed)` (backend/model.py:59-62), like gru_fwd.cu,
// but moves gh_t = h_{t-1} W_hh^T from fp32 FMAs + warp shuffles to the tensor cores while keeping
// fp32-level accuracy:
//
//   * a thread-block cluster of 8 CTAs owns one tile of 128 length-sorted rows and one direction
//     for ALL timesteps.  CTA c owns hidden units [32c, 32c+32), i.e. 96 gate columns (r, z, n).
//   * its slice of W_hh lives in shared memory for the whole kernel as the UMMA B operand, split
//     into two fp16 planes  W = W_hi + W_lo  (2 x 48 KB);  h_{t-1} of the whole tile is the A
//     operand, split the same way (2 x 64 KB).  Three kind::f16 MMA chains per step,
//     h_hi W_hi + h_hi W_lo + h_lo W_hi, accumulate in fp32 in tensor memory: the dropped
//     h_lo W_lo term is 2^-22 relative (W in [-1/16, 1/16] and h in [-1, 1] keep both planes in
//     fp16 range; the low planes use fp16 subnormals, which the tensor core handles exactly).
//   * epilogue thread = (row, 16 units): tcgen05.ld of its 48 gate pre-activations, fused gate
//     math against the gi row segment it prefetched while the MMAs ran, fp32 h kept in registers
//     across timesteps, results written to y / saved / h_last, and the new h values written as
//     fp16 hi/lo straight into the CTA's own slice of the A operand.
//   * exchange: one elected thread pushes that 2 x 8 KB slice into the other seven CTAs' A
//     operands with cp.async.bulk shared::cta -> shared::cluster copies that complete on the
//     DESTINATION's mbarrier — no cluster-wide barrier in the loop.  The write-after-read hazard
//     (a peer's MMA still reading its operand) is covered by a second mbarrier that every CTA's
//     MMA completion signals in all eight CTAs (tcgen05.commit ... multicast::cluster).
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

constexpr int TC_H = 256;
constexpr int TC_CL = 8;                         // CTAs per cluster
constexpr int TC_UN = TC_H / TC_CL;              // 32 hidden units per CTA
constexpr int TC_NG = 3 * TC_UN;                 // 96 gate columns per CTA (UMMA N)
constexpr int TC_ROWS = 128;                     // rows per tile (UMMA M)
constexpr int TC_KC = TC_H / 8;                  // 32 k-chunks of 8 halves
constexpr int TC_A_LBO = TC_ROWS * 16;           // 2048 B between k-chunks of the A operand
constexpr int TC_B_LBO = TC_NG * 16;             // 1536 B between k-chunks of the B operand
constexpr int TC_A_BYTES = TC_KC * TC_A_LBO;     // 64 KB per plane
constexpr int TC_B_BYTES = TC_KC * TC_B_LBO;     // 48 KB per plane
constexpr int TC_SLICE_BYTES = (TC_UN / 8) * TC_A_LBO;   // 8 KB: one CTA's units in one plane
constexpr int TC_THREADS = 256;
constexpr int TC_TMEM_COLS = 128;
constexpr int TC_SMEM = 2 * TC_A_BYTES + 2 * TC_B_BYTES + TC_NG * 4 + 3 * TC_ROWS * 4 + 3 * 8 + 8;

struct GruTcArgs {
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

namespace {

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cta -> (remote) shared::cluster bulk copy, completion counted in bytes on the destination's mbarrier
__device__ __forceinline__ void bulk_s2s(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes, uint32_t bar_cluster) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   dst_cluster),
               "r"(src_cta), "r"(bytes), "r"(bar_cluster)
               : "memory");
}
__device__ __forceinline__ void mma_commit_multicast(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          ptx::smem_u32(bar)),
      "h"(mask)
      : "memory");
}
__device__ __forceinline__ void mma_f16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
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
// kind::f16 instruction descriptor: fp16 A and B (format 0), fp32 accumulate, both K-major, dense
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// x = hi + lo with both halves in fp16 (lo may be subnormal): packs 8 values into two 16-byte rows
__device__ __forceinline__ void split8(const float* x, uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __half h0 = __float2half_rn(x[2 * i]), h1 = __float2half_rn(x[2 * i + 1]);
    const __half l0 = __float2half_rn(x[2 * i] - __half2float(h0));
    const __half l1 = __float2half_rn(x[2 * i + 1] - __half2float(h1));
    h[i] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
    l[i] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float fast_tanh(float x) { return 1.0f - __fdividef(2.0f, 1.0f + __expf(2.0f * x)); }

}  // namespace

__global__ void __cluster_dims__(TC_CL, 1, 1) __launch_bounds__(TC_THREADS, 1)
gru_fwd_tc_kernel(GruTcArgs a) {
  extern __shared__ __align__(128) unsigned char sm[];
  unsigned char* h_hi = sm;
  unsigned char* h_lo = sm + TC_A_BYTES;
  unsigned char* w_hi = sm + 2 * TC_A_BYTES;
  unsigned char* w_lo = w_hi + TC_B_BYTES;
  float* bias = reinterpret_cast<float*>(w_lo + TC_B_BYTES);      // [96]: b_hr, b_hz, b_hn of the slice
  int* lens = reinterpret_cast<int*>(bias + TC_NG);
  int* toff = lens + TC_ROWS;
  int* rowid = toff + TC_ROWS;
  uint64_t* h_full = reinterpret_cast<uint64_t*>(rowid + TC_ROWS);   // peers' slices have landed
  uint64_t* mma_done = h_full + 1;                                   // own accumulators are ready
  uint64_t* consumed = h_full + 2;                                   // all 8 CTAs finished reading h_{t-1}
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(h_full + 3);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rank = (int)cluster_ctarank();
  const int tile = blockIdx.x / TC_CL;
  const int dir = blockIdx.y;
  const int G3 = 3 * TC_H;
  const int s0 = tile * TC_ROWS;

  if (tid == 0) {
    ptx::mbar_init(h_full, 1);
    ptx::mbar_init(mma_done, 1);
    ptx::mbar_init(consumed, TC_CL);
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, TC_TMEM_COLS);
    ptx::tmem_relinquish();
  }
  for (int i = tid; i < TC_ROWS; i += TC_THREADS) {
    const int s = s0 + i;
    if (s < a.B) {
      const int off = a.offsets[s];
      lens[i] = a.offsets[s + 1] - off;
      toff[i] = off;
      rowid[i] = a.order[s];
    } else {
      lens[i] = 0; toff[i] = 0; rowid[i] = 0;
    }
  }
  {  // h_0 = 0 in both planes
    uint4* p = reinterpret_cast<uint4*>(h_hi);
    for (int i = tid; i < 2 * TC_A_BYTES / 16; i += TC_THREADS) p[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  {  // resident weights: rows (gate g, unit 32*rank + u) of W_hh[dir], split into fp16 planes
    const float* wbase = a.w_hh + (size_t)dir * G3 * TC_H;
    for (int idx = tid; idx < TC_NG * TC_KC; idx += TC_THREADS) {
      const int n = idx % TC_NG, kc = idx / TC_NG;
      const int g = n / TC_UN, u = n % TC_UN;
      const float4* src = reinterpret_cast<const float4*>(wbase + (size_t)(g * TC_H + rank * TC_UN + u) * TC_H + kc * 8);
      const float4 v0 = __ldg(src), v1 = __ldg(src + 1);
      const float x[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
      uint4 hi, lo;
      split8(x, hi, lo);
      *reinterpret_cast<uint4*>(w_hi + kc * TC_B_LBO + n * 16) = hi;
      *reinterpret_cast<uint4*>(w_lo + kc * TC_B_LBO + n * 16) = lo;
    }
    if (tid < TC_NG) bias[tid] = a.b_hh[dir * G3 + (tid / TC_UN) * TC_H + rank * TC_UN + (tid % TC_UN)];
  }
  ptx::fence_proxy_async_smem();          // operand bytes written by threads -> visible to the tensor core
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  cluster_sync_all();                     // every CTA's barriers are initialised before any remote arrive
  const uint32_t tmem_base = *tmem_slot;

  // epilogue role of this thread: TMEM lane quarter q (rows 32q..32q+31), unit half uh
  const int q = warp & 3, uh = warp >> 2;
  const int row = q * 32 + lane;
  const int u0 = uh * 16;                           // first owned unit inside the CTA's slice
  const int j0 = rank * TC_UN + u0;                 // ... as a global hidden unit
  const uint32_t tmem_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)u0;
  const int len = lens[row];
  const int tbase = toff[row];
  const int maxlen = lens[0];
  const int gi_ld = a.dirs * G3, y_ld = a.dirs * TC_H;
  const float* gi_base = a.gi + dir * G3 + j0;
  float h[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) h[i] = 0.f;

  const uint32_t h_hi_u32 = ptx::smem_u32(h_hi), h_lo_u32 = ptx::smem_u32(h_lo);
  const uint32_t w_hi_u32 = ptx::smem_u32(w_hi), w_lo_u32 = ptx::smem_u32(w_lo);
  constexpr uint32_t idesc = make_idesc_f16(TC_ROWS, TC_NG);

  for (int t = 0; t < maxlen; ++t) {
    const uint32_t par = (uint32_t)t & 1u;
    const bool active = t < len;
    const int tok = active ? tbase + (dir == 0 ? t : len - 1 - t) : 0;

    if (warp == 0) {
      // ===== MMA issue: gh = h_hi W_hi + h_hi W_lo + h_lo W_hi (16 k-steps of 16 each) =====
      if (t > 0) ptx::mbar_wait(h_full, par ^ 1u);
      ptx::tc_fence_after_sync();
      if (ptx::elect_one()) {
        const uint64_t a_hi = make_kmajor_nosw_desc(h_hi_u32, TC_A_LBO, 128);
        const uint64_t a_lo = make_kmajor_nosw_desc(h_lo_u32, TC_A_LBO, 128);
        const uint64_t b_hi = make_kmajor_nosw_desc(w_hi_u32, TC_B_LBO, 128);
        const uint64_t b_lo = make_kmajor_nosw_desc(w_lo_u32, TC_B_LBO, 128);
#pragma unroll
        for (int ks = 0; ks < TC_H / 16; ++ks)
          mma_f16_ss(tmem_base, a_hi + (uint64_t)(ks * (2 * TC_A_LBO >> 4)), b_hi + (uint64_t)(ks * (2 * TC_B_LBO >> 4)),
                     idesc, ks != 0);
#pragma unroll
        for (int ks = 0; ks < TC_H / 16; ++ks)
          mma_f16_ss(tmem_base, a_hi + (uint64_t)(ks * (2 * TC_A_LBO >> 4)), b_lo + (uint64_t)(ks * (2 * TC_B_LBO >> 4)),
                     idesc, 1u);
#pragma unroll
        for (int ks = 0; ks < TC_H / 16; ++ks)
          mma_f16_ss(tmem_base, a_lo + (uint64_t)(ks * (2 * TC_A_LBO >> 4)), b_hi + (uint64_t)(ks * (2 * TC_B_LBO >> 4)),
                     idesc, 1u);
        ptx::mma_commit(mma_done);
        // nobody waits for the last step's signal, and a peer may have left by the time it would land
        if (t + 1 < maxlen) mma_commit_multicast(consumed, (uint16_t)0xff);
      }
      __syncwarp();
    }

    // gi row segment of this step (3 gates x 16 units), in flight while the MMAs run
    float4 g4[12];
    if (active) {
      const float4* gp = reinterpret_cast<const float4*>(gi_base + (size_t)tok * gi_ld);
#pragma unroll
      for (int g = 0; g < 3; ++g)
#pragma unroll
        for (int i = 0; i < 4; ++i) g4[g * 4 + i] = __ldg(gp + g * (TC_H / 4) + i);
    }

    ptx::mbar_wait(mma_done, par);
    ptx::tc_fence_after_sync();
    if (__any_sync(0xffffffffu, active)) {
      uint32_t ar[16], az[16], an[16];
      tmem_ld_32x16(tmem_row, ar);
      tmem_ld_32x16(tmem_row + TC_UN, az);
      tmem_ld_32x16(tmem_row + 2 * TC_UN, an);
      ptx::tmem_ld_wait();
      if (active) {
        const float* gf = reinterpret_cast<const float*>(g4);
        float rr[16], zz[16], nn[16], gn[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float r = fast_sigmoid(gf[i] + __uint_as_float(ar[i]) + bias[u0 + i]);
          const float z = fast_sigmoid(gf[16 + i] + __uint_as_float(az[i]) + bias[TC_UN + u0 + i]);
          const float ghn = __uint_as_float(an[i]) + bias[2 * TC_UN + u0 + i];
          const float n = fast_tanh(fmaf(r, ghn, gf[32 + i]));
          h[i] = fmaf(z, h[i] - n, n);            // (1-z)*n + z*h
          rr[i] = r; zz[i] = z; nn[i] = n; gn[i] = ghn;
        }
        // new h -> own slice of the A operand (k = 32*rank + u0 + i -> k-chunks 4*rank + 2*uh + {0,1})
        {
          uint4 hi, lo;
          const int kc = 4 * rank + 2 * uh;
          split8(h, hi, lo);
          *reinterpret_cast<uint4*>(h_hi + kc * TC_A_LBO + row * 16) = hi;
          *reinterpret_cast<uint4*>(h_lo + kc * TC_A_LBO + row * 16) = lo;
          split8(h + 8, hi, lo);
          *reinterpret_cast<uint4*>(h_hi + (kc + 1) * TC_A_LBO + row * 16) = hi;
          *reinterpret_cast<uint4*>(h_lo + (kc + 1) * TC_A_LBO + row * 16) = lo;
        }
        if (a.y) {
          float4* yp = reinterpret_cast<float4*>(a.y + (size_t)tok * y_ld + dir * TC_H + j0);
#pragma unroll
          for (int i = 0; i < 4; ++i) yp[i] = make_float4(h[4 * i], h[4 * i + 1], h[4 * i + 2], h[4 * i + 3]);
        }
        if (a.saved) {
          float4* sv = reinterpret_cast<float4*>(a.saved + ((size_t)tok * a.dirs + dir) * 4 * TC_H + j0);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            sv[i] = make_float4(rr[4 * i], rr[4 * i + 1], rr[4 * i + 2], rr[4 * i + 3]);
            sv[TC_H / 4 + i] = make_float4(zz[4 * i], zz[4 * i + 1], zz[4 * i + 2], zz[4 * i + 3]);
            sv[2 * TC_H / 4 + i] = make_float4(nn[4 * i], nn[4 * i + 1], nn[4 * i + 2], nn[4 * i + 3]);
            sv[3 * TC_H / 4 + i] = make_float4(gn[4 * i], gn[4 * i + 1], gn[4 * i + 2], gn[4 * i + 3]);
          }
        }
        if (t == len - 1) {
          float4* hp = reinterpret_cast<float4*>(a.h_last + (size_t)rowid[row] * y_ld + dir * TC_H + j0);
#pragma unroll
          for (int i = 0; i < 4; ++i) hp[i] = make_float4(h[4 * i], h[4 * i + 1], h[4 * i + 2], h[4 * i + 3]);
        }
      }
    }
    ptx::tc_fence_before_sync();
    ptx::fence_proxy_async_smem();        // the slice written above is read by the copy engine and the tensor core
    __syncthreads();

    if (warp == 0 && t + 1 < maxlen) {
      // ===== exchange: push the slice to the seven peers once all of them have finished reading h_{t-1} =====
      if (ptx::elect_one()) {
        ptx::mbar_wait(consumed, par);
        ptx::mbar_arrive_expect_tx(h_full, (TC_CL - 1) * 2 * TC_SLICE_BYTES);
        const uint32_t off = (uint32_t)rank * TC_SLICE_BYTES;
        const uint32_t bar = ptx::smem_u32(h_full);
#pragma unroll
        for (int p = 0; p < TC_CL; ++p) {
          if (p == rank) continue;
          const uint32_t rbar = mapa(bar, (uint32_t)p);
          bulk_s2s(mapa(h_hi_u32 + off, (uint32_t)p), h_hi_u32 + off, TC_SLICE_BYTES, rbar);
          bulk_s2s(mapa(h_lo_u32 + off, (uint32_t)p), h_lo_u32 + off, TC_SLICE_BYTES, rbar);
        }
      }
      __syncwarp();
    }
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();                     // no CTA leaves while a peer's commit may still arrive on its barriers
  if (warp == 1) ptx::tmem_dealloc(tmem_base, TC_TMEM_COLS);
}

int launch_gru_fwd_tc(const float* gi, const float* w_hh, const float* b_hh, const int32_t* order,
                      const int32_t* offsets, int B, int dirs, float* y, float* h_last, float* saved, cudaStream_t st) {
  GruTcArgs a{gi, w_hh, b_hh, order, offsets, B, dirs, y, h_last, saved};
  TTR_CHECK_CUDA(cudaFuncSetAttribute(gru_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM));
  dim3 grid(ceil_div(B, TC_ROWS) * TC_CL, dirs);
  gru_fwd_tc_kernel<<<grid, TC_THREADS, TC_SMEM, st>>>(a);
  TTR_CHECK_LAUNCH();
  return TTR_OK;
}

}  // namespace ttr
