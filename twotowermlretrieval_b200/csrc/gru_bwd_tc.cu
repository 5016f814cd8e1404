// gru_bwd_tc.cu — backpropagation through time of one GRU layer (H = 256) with the recurrent product
// dh_{t-1} += d(gh)_t W_hh on tcgen05.
//
// The autograd of `self.rnn(packed)` (backend/model.py:59-62) reached from `loss.backward()`
// (backend/main.py:254), like gru_bwd.cu.  The forward kernel (gru_fwd_tc.cu) splits the OUTPUT units of
// gh = h W_hh^T over the 8 CTAs of a cluster and all-gathers h; the backward product reduces over all 768
// gate rows, so the same unit ownership makes it a split-K problem instead:
//
//   * cluster of 8 CTAs per (128-row tile, direction); CTA c owns hidden units [32c, 32c+32) and therefore
//     the 96 gate rows (r, z, n of those units) of d(gh) — which it computes itself, elementwise, from
//     its own slice of dh.  No exchange is needed BEFORE the product.
//   * B operand, resident for the whole kernel: W_hh[own 96 gate rows, all 256 columns] as fp16
//     ([N = 256][K = 96], 48 KB).  A operand, rewritten every step: d(gh)[128 rows, own 96] * 2^14 as two
//     fp16 planes (hi + lo: 22 significant bits; the power-of-two scale keeps gradients down to ~4e-9 in
//     the normal range).  12 kind::f16 MMAs (M128 N256 K16) per step give the CTA's PARTIAL
//     dh_{t-1}[128, 256] in tensor memory.
//   * reduce-scatter through L2: the partial goes TMEM -> swizzled smem staging -> cp.reduce.async.bulk.tensor
//     (.add.f32) into a per-cluster accumulator image [2 parities][128][256] fp32; the copy engine and the L2
//     do the additions.  When its reductions have completed a CTA signals an mbarrier in all 8 CTAs (remote
//     arrive); a CTA that has collected all 16 signals reads its own 32 columns of the sum, zeroes them for
//     the step after next, and continues with the elementwise gate gradients.
//   * the elementwise phase is not tied to tensor-memory lanes, so threads are mapped (row, 4 units) with
//     8 lanes covering a row's 128-byte segment: every global load/store of saved gates, dy, h_prev, d(gi),
//     d(gh) and the accumulator is coalesced without shuffles.
//   * the bias gradients (column sums of d(gi) and d(gh) over all tokens: what `ttr_colsum` would re-read
//     12 KB per token for) are accumulated in registers on the way and added to db_ih / db_hh once per CTA.
#include <cudaTypedefs.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "ptx.cuh"

namespace ttr {

extern int g_debug_flags;
int make_rowmajor_map(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int box_rows, bool tf32);

constexpr int BT_H = 256;
constexpr int BT_CL = 8;
constexpr int BT_UN = BT_H / BT_CL;            // 32 units per CTA
constexpr int BT_K = 3 * BT_UN;                // 96 gate rows per CTA = K of the product
constexpr int BT_ROWS = 128;                   // rows per tile (UMMA M)
constexpr int BT_KC = BT_K / 8;                // 12 k-chunks of 8 halves
constexpr int BT_A_LBO = BT_ROWS * 16;         // 2048
constexpr int BT_B_LBO = BT_H * 16;            // 4096
constexpr int BT_A_BYTES = BT_KC * BT_A_LBO;   // 24 KB per plane
constexpr int BT_B_BYTES = BT_KC * BT_B_LBO;   // 48 KB
constexpr int BT_STG_BYTES = BT_ROWS * 128;    // one 128-row x 32-float chunk, SWIZZLE_128B
constexpr int BT_THREADS = 256;
constexpr int BT_TMEM_COLS = 256;
constexpr float BT_SCALE = 16384.0f;
constexpr int BT_SMEM = 2 * BT_A_BYTES + BT_B_BYTES + 4 * BT_STG_BYTES + 3 * BT_ROWS * 4 + 2 * 8 + 16 + 1024;
constexpr int64_t BT_ACC_ROWS_PER_CLUSTER = 2 * BT_ROWS;     // two parities

struct GruBwdTcArgs {
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
  float* acc;             // [clusters][2][128][256] fp32, zero on entry
  float* db_ih;           // [dirs*3H] += column sums of d(gi)  (or null)
  float* db_hh;           // [dirs*3H] += column sums of d(gh)  (or null)
};

namespace {

__device__ __forceinline__ uint32_t cluster_ctarank_b() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa_b(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void remote_arrive(uint32_t bar_cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}
__device__ __forceinline__ void cluster_sync_b() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d_f32(const CUtensorMap* map, const void* smem_src, int32_t c0, int32_t c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map),
               "r"(ptx::smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ uint64_t nosw_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// 4 scaled values -> 4 fp16 hi (8 bytes) and 4 fp16 lo (8 bytes)
__device__ __forceinline__ void split4(const float (&xin)[4], uint2& hi, uint2& lo) {
  float x[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) x[e] = fminf(fmaxf(xin[e], -60000.f), 60000.f);   // |d(gh)| > 3.6 saturates instead of overflowing
  const __half2 h0 = __floats2half2_rn(x[0], x[1]), h1 = __floats2half2_rn(x[2], x[3]);
  const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
  const __half2 l0 = __floats2half2_rn(x[0] - f0.x, x[1] - f0.y), l1 = __floats2half2_rn(x[2] - f1.x, x[3] - f1.y);
  hi = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
  lo = make_uint2(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1));
}

}  // namespace

__global__ void __cluster_dims__(BT_CL, 1, 1) __launch_bounds__(BT_THREADS, 1)
gru_bwd_tc_kernel(GruBwdTcArgs a, const __grid_constant__ CUtensorMap map_acc) {
  extern __shared__ unsigned char smem_raw_b[];
  unsigned char* base = smem_raw_b + ((1024u - (ptx::smem_u32(smem_raw_b) & 1023u)) & 1023u);
  unsigned char* stg = base;                                   // [2 halves][2 buffers][16 KB], 1024-aligned
  unsigned char* a_hi = stg + 4 * BT_STG_BYTES;
  unsigned char* a_lo = a_hi + BT_A_BYTES;
  unsigned char* w_sm = a_lo + BT_A_BYTES;
  int* lens = reinterpret_cast<int*>(w_sm + BT_B_BYTES);
  int* toff = lens + BT_ROWS;
  int* rowid = toff + BT_ROWS;
  uint64_t* mma_done = reinterpret_cast<uint64_t*>(rowid + BT_ROWS);
  uint64_t* red_done = mma_done + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(red_done + 1);

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int rank = (int)cluster_ctarank_b();
  const int tile = blockIdx.x / BT_CL;
  const int dir = blockIdx.y;
  const int G3 = 3 * BT_H;
  const int s0 = tile * BT_ROWS;
  const int cid = blockIdx.y * (gridDim.x / BT_CL) + tile;

  if (tid == 0) {
    ptx::mbar_init(mma_done, 1);
    ptx::mbar_init(red_done, 2 * BT_CL);       // every CTA's two column halves signal
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, BT_TMEM_COLS);
    ptx::tmem_relinquish();
  }
  for (int i = tid; i < BT_ROWS; i += BT_THREADS) {
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
  {  // B operand: element (n, k) = W_hh[dir][g*256 + 32*rank + u][n], k = 32*g + u; layout [k/8][n][8 halves]
    const float* wbase = a.w_hh + (size_t)dir * G3 * BT_H;
    for (int idx = tid; idx < BT_H * BT_KC; idx += BT_THREADS) {
      const int n = idx % BT_H, kc = idx / BT_H;
      uint32_t pk[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int k0 = kc * 8 + 2 * i, k1 = k0 + 1;
        const float v0 = __ldg(wbase + (size_t)((k0 / BT_UN) * BT_H + rank * BT_UN + (k0 % BT_UN)) * BT_H + n);
        const float v1 = __ldg(wbase + (size_t)((k1 / BT_UN) * BT_H + rank * BT_UN + (k1 % BT_UN)) * BT_H + n);
        const __half2 h = __floats2half2_rn(v0, v1);
        pk[i] = *reinterpret_cast<const uint32_t*>(&h);
      }
      *reinterpret_cast<uint4*>(w_sm + kc * BT_B_LBO + n * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
  }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  cluster_sync_b();
  const uint32_t tmem_base = *tmem_slot;

  const int maxlen = lens[0];
  const int g_ld = a.dirs * G3, y_ld = a.dirs * BT_H;
  // elementwise phase: thread = (rows rsub + 32*it, units 4*c4 .. 4*c4+3 of the CTA's slice)
  const int rsub = tid >> 3, c4 = tid & 7;
  const int j4 = rank * BT_UN + 4 * c4;                       // first of the thread's 4 global hidden units
  float dbs[4][4];                                            // running sums of dr, dz, dn, dn*r over rows and steps
#pragma unroll
  for (int g = 0; g < 4; ++g)
#pragma unroll
    for (int e = 0; e < 4; ++e) dbs[g][e] = 0.f;
  float dhz[4][4];                                            // dh_t * z carried to the next step
#pragma unroll
  for (int it = 0; it < 4; ++it)
#pragma unroll
    for (int e = 0; e < 4; ++e) dhz[it][e] = 0.f;
  float* acc_cluster = a.acc + (size_t)cid * BT_ACC_ROWS_PER_CLUSTER * BT_H;
  // Inputs of a step that do not depend on the recurrence (saved gates, dy, h_prev) are fetched one step ahead,
  // right after the MMAs are issued: their HBM latency hides behind the MMA, the reduce and the cluster signal
  // instead of heading the next step's dependency chain (the step is a latency chain: with 16 clusters in flight
  // at 1,024 rows nothing else fills the SM).
  float4 pre[4][6];                                           // [item][r, z, n, ghn, dy, h_prev]
  auto prefetch = [&](int s) {
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int row = rsub + 32 * it;
      const int len = lens[row];
#pragma unroll
      for (int v = 0; v < 6; ++v) pre[it][v] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (s >= 0 && s < len) {
        const int pos = (dir == 0) ? s : len - 1 - s;
        const int tok = toff[row] + pos;
        const float* sv = a.saved + ((size_t)tok * a.dirs + dir) * 4 * BT_H + j4;
#pragma unroll
        for (int v = 0; v < 4; ++v) pre[it][v] = __ldg(reinterpret_cast<const float4*>(sv + v * BT_H));
        if (a.dy) pre[it][4] = __ldg(reinterpret_cast<const float4*>(a.dy + (size_t)tok * y_ld + dir * BT_H + j4));
        if (s > 0) {
          const int tokp = toff[row] + ((dir == 0) ? s - 1 : len - s);
          pre[it][5] = __ldg(reinterpret_cast<const float4*>(a.y + (size_t)tokp * y_ld + dir * BT_H + j4));
        }
      }
    }
  };
  prefetch(maxlen - 1);
  // reduce phase: warp = (TMEM lane quarter q, column half uh)
  const int q = warp & 3, uh = warp >> 2;
  const int r_in_tile = q * 32 + lane;
  unsigned char* my_stg = stg + uh * 2 * BT_STG_BYTES;
  const uint32_t a_hi_u32 = ptx::smem_u32(a_hi), a_lo_u32 = ptx::smem_u32(a_lo), w_u32 = ptx::smem_u32(w_sm);
  constexpr uint32_t idesc = ptx::make_idesc_f16(BT_ROWS, BT_H);
  int chunk_no = 0;

  for (int s = maxlen - 1, step = 0; s >= 0; --s, ++step) {
    const uint32_t par = (uint32_t)step & 1u;
    const int pbuf_read = (s + 1) & 1, pbuf_write = s & 1;
    if (step > 0) {
      ptx::mbar_wait(red_done, par ^ 1u);                     // every CTA's partial of step s+1 has been added
      __threadfence();
    }
    // ---------------- elementwise gate gradients of the CTA's own units ----------------
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int row = rsub + 32 * it;
      const int len = lens[row];
      float dgh_s[3][4];                                      // scaled d(gh) r, z, n for the A operand
#pragma unroll
      for (int g = 0; g < 3; ++g)
#pragma unroll
        for (int e = 0; e < 4; ++e) dgh_s[g][e] = 0.f;
      if (s < len) {
        const int pos = (dir == 0) ? s : len - 1 - s;
        const int tok = toff[row] + pos;
        float4 dh4 = make_float4(dhz[it][0], dhz[it][1], dhz[it][2], dhz[it][3]);
        if (s < len - 1) {
          float4* ap = reinterpret_cast<float4*>(acc_cluster + ((size_t)pbuf_read * BT_ROWS + row) * BT_H + j4);
          const float4 av = __ldcg(ap);
          *ap = make_float4(0.f, 0.f, 0.f, 0.f);              // ready for the step after next
          dh4.x = fmaf(av.x, 1.0f / BT_SCALE, dh4.x); dh4.y = fmaf(av.y, 1.0f / BT_SCALE, dh4.y);
          dh4.z = fmaf(av.z, 1.0f / BT_SCALE, dh4.z); dh4.w = fmaf(av.w, 1.0f / BT_SCALE, dh4.w);
        }
        {
          const float4 v = pre[it][4];
          dh4.x += v.x; dh4.y += v.y; dh4.z += v.z; dh4.w += v.w;
        }
        if (s == len - 1 && a.dh_last) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(a.dh_last + (size_t)rowid[row] * y_ld + dir * BT_H + j4));
          dh4.x += v.x; dh4.y += v.y; dh4.z += v.z; dh4.w += v.w;
        }
        const float4 r4 = pre[it][0], z4 = pre[it][1], n4 = pre[it][2], g4 = pre[it][3], hp = pre[it][5];
        const float dh[4] = {dh4.x, dh4.y, dh4.z, dh4.w};
        const float r[4] = {r4.x, r4.y, r4.z, r4.w}, z[4] = {z4.x, z4.y, z4.z, z4.w};
        const float n[4] = {n4.x, n4.y, n4.z, n4.w}, ghn[4] = {g4.x, g4.y, g4.z, g4.w};
        const float hpv[4] = {hp.x, hp.y, hp.z, hp.w};
        float dr[4], dz[4], dn[4], dgn[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          dn[e] = dh[e] * (1.f - z[e]) * (1.f - n[e] * n[e]);
          dz[e] = dh[e] * (hpv[e] - n[e]) * z[e] * (1.f - z[e]);
          dr[e] = dn[e] * ghn[e] * r[e] * (1.f - r[e]);
          dgn[e] = dn[e] * r[e];
          dhz[it][e] = dh[e] * z[e];
          dbs[0][e] += dr[e]; dbs[1][e] += dz[e]; dbs[2][e] += dn[e]; dbs[3][e] += dgn[e];
          dgh_s[0][e] = dr[e] * BT_SCALE; dgh_s[1][e] = dz[e] * BT_SCALE; dgh_s[2][e] = dgn[e] * BT_SCALE;
        }
        float* gi = a.dgi + (size_t)tok * g_ld + dir * G3 + j4;
        *reinterpret_cast<float4*>(gi) = make_float4(dr[0], dr[1], dr[2], dr[3]);
        *reinterpret_cast<float4*>(gi + BT_H) = make_float4(dz[0], dz[1], dz[2], dz[3]);
        *reinterpret_cast<float4*>(gi + 2 * BT_H) = make_float4(dn[0], dn[1], dn[2], dn[3]);
        float* gh = a.dgh + (size_t)tok * g_ld + dir * G3 + j4;
        *reinterpret_cast<float4*>(gh) = make_float4(dr[0], dr[1], dr[2], dr[3]);
        *reinterpret_cast<float4*>(gh + BT_H) = make_float4(dz[0], dz[1], dz[2], dz[3]);
        *reinterpret_cast<float4*>(gh + 2 * BT_H) = make_float4(dgn[0], dgn[1], dgn[2], dgn[3]);
      }
      // A operand rows (zeros for rows that are not active: their partial must stay exactly zero)
#pragma unroll
      for (int g = 0; g < 3; ++g) {
        uint2 hi, lo;
        split4(dgh_s[g], hi, lo);
        const int kc = g * 4 + (c4 >> 1);                     // k = 32 g + 4 c4 + e
        const int off = kc * BT_A_LBO + row * 16 + (c4 & 1) * 8;
        *reinterpret_cast<uint2*>(a_hi + off) = hi;
        *reinterpret_cast<uint2*>(a_lo + off) = lo;
      }
    }
    if (s == 0) break;                                        // nobody needs dh_{-1}
    asm volatile("fence.proxy.async;" ::: "memory");         // A operand (smem) and the zeroed accumulator (global)
    ptx::tc_fence_before_sync();
    __syncthreads();
    ptx::tc_fence_after_sync();

    // ---------------- partial dh_{s-1} = d(gh)[:, own 96] * W_hh[own 96, :] ----------------
    if (warp == 0) {
      if (ptx::elect_one()) {
        const uint64_t ahi = nosw_desc(a_hi_u32, BT_A_LBO, 128), alo = nosw_desc(a_lo_u32, BT_A_LBO, 128);
        const uint64_t bd = nosw_desc(w_u32, BT_B_LBO, 128);
#pragma unroll
        for (int ks = 0; ks < BT_K / 16; ++ks)
          ptx::mma_f16_ss(tmem_base, ahi + (uint64_t)(ks * (2 * BT_A_LBO >> 4)), bd + (uint64_t)(ks * (2 * BT_B_LBO >> 4)), idesc,
                          ks != 0);
#pragma unroll
        for (int ks = 0; ks < BT_K / 16; ++ks)
          ptx::mma_f16_ss(tmem_base, alo + (uint64_t)(ks * (2 * BT_A_LBO >> 4)), bd + (uint64_t)(ks * (2 * BT_B_LBO >> 4)), idesc,
                          1u);
        ptx::mma_commit(mma_done);
      }
      __syncwarp();
    }
    prefetch(s - 1);                                          // next step's inputs, in flight during MMA + reduce
    ptx::mbar_wait(mma_done, par);
    ptx::tc_fence_after_sync();

    // ---------------- reduce-scatter: partial -> accumulator image in L2 (copy-engine adds) ----------------
    const int acc_row0 = (int)((int64_t)cid * BT_ACC_ROWS_PER_CLUSTER) + pbuf_write * BT_ROWS;
#pragma unroll 1
    for (int c = 0; c < 4; ++c, ++chunk_no) {
      uint32_t rr[32];
      ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + uh * 128 + c * 32, rr);
      ptx::tmem_ld_wait();
      unsigned char* sb = my_stg + (chunk_no & 1) * BT_STG_BYTES;
      if (q == 0 && lane == 0) ptx::bulk_wait_group_read<1>();          // the reduce that used this buffer two chunks ago
      ptx::named_bar_sync(1 + uh, 128);
      uint4* rowp = reinterpret_cast<uint4*>(sb + r_in_tile * 128);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        rowp[j ^ (r_in_tile & 7)] = make_uint4(rr[4 * j], rr[4 * j + 1], rr[4 * j + 2], rr[4 * j + 3]);
      ptx::fence_proxy_async_smem();
      ptx::named_bar_sync(1 + uh, 128);
      if (q == 0 && lane == 0) {
        tma_reduce_add_2d_f32(&map_acc, sb, uh * 128 + c * 32, acc_row0);
        ptx::bulk_commit_group();
      }
    }
    ptx::tc_fence_before_sync();
    if (q == 0 && lane == 0) {
      ptx::bulk_wait_group<0>();                              // this half's additions have been performed
      __threadfence();
      const uint32_t bar = ptx::smem_u32(red_done);
#pragma unroll
      for (int p = 0; p < BT_CL; ++p) remote_arrive(mapa_b(bar, (uint32_t)p));
    }
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  if (a.db_ih || a.db_hh) {
    // bias gradients: reduce the 32 row groups through shared memory (the staging area is idle now), one global add per column
    float* red = reinterpret_cast<float*>(stg);               // [4 arrays][32 units]
    if (tid < 4 * BT_UN) red[tid] = 0.f;
    __syncthreads();
#pragma unroll
    for (int g = 0; g < 4; ++g)
#pragma unroll
      for (int e = 0; e < 4; ++e) atomicAdd(&red[g * BT_UN + 4 * c4 + e], dbs[g][e]);
    __syncthreads();
    if (tid < 3 * BT_UN) {
      const int g = tid / BT_UN, u = tid % BT_UN;
      const int col = dir * G3 + g * BT_H + rank * BT_UN + u;
      if (a.db_ih) atomicAdd(a.db_ih + col, red[g * BT_UN + u]);                       // dr, dz, dn
      if (a.db_hh) atomicAdd(a.db_hh + col, red[(g == 2 ? 3 : g) * BT_UN + u]);        // dr, dz, dn * r
    }
  }
  cluster_sync_b();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, BT_TMEM_COLS);
}

int64_t gru_bwd_tc_workspace_bytes(int B, int dirs) {
  return (int64_t)ceil_div(B, BT_ROWS) * dirs * BT_ACC_ROWS_PER_CLUSTER * BT_H * 4;
}

int launch_gru_bwd_tc(const float* dy, const float* dh_last, const float* y, const float* saved, const float* w_hh,
                      const int32_t* order, const int32_t* offsets, int B, int dirs, float* dgi, float* dgh,
                      void* workspace, float* db_ih, float* db_hh, cudaStream_t st) {
  const int64_t ws_bytes = gru_bwd_tc_workspace_bytes(B, dirs);
  TTR_CHECK_CUDA(cudaMemsetAsync(workspace, 0, (size_t)ws_bytes, st));
  CUtensorMap map_acc;
  const int64_t rows = (int64_t)ceil_div(B, BT_ROWS) * dirs * BT_ACC_ROWS_PER_CLUSTER;
  int rc = make_rowmajor_map(&map_acc, reinterpret_cast<const float*>(workspace), rows, BT_H, BT_ROWS, false);
  if (rc != TTR_OK) return rc;
  GruBwdTcArgs a{dy, dh_last, y, saved, w_hh, order, offsets, B, dirs, dgi, dgh, reinterpret_cast<float*>(workspace),
                 db_ih, db_hh};
  TTR_CHECK_CUDA(cudaFuncSetAttribute(gru_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BT_SMEM));
  dim3 grid(ceil_div(B, BT_ROWS) * BT_CL, dirs);
  gru_bwd_tc_kernel<<<grid, BT_THREADS, BT_SMEM, st>>>(a, map_acc);
  TTR_CHECK_LAUNCH();
  return TTR_OK;
}

}  // namespace ttr
