// gemm_tcgen05.cu — GRU input projection C[M,N] = A[M,K] * W[N,K]^T + bias on the 5th-gen
// tensor cores: TMA (SWIZZLE_128B) -> shared-memory ring -> tcgen05.mma kind::tf32 with the
// fp32 accumulator in TMEM (double-buffered) -> tcgen05.ld epilogue with fused bias.
//
// Replaces the `X W_ih^T + b_ih` half of nn.GRU (backend/model.py:59-62) for every timestep
// and both directions in one launch (N = dirs*3H).  Persistent: one CTA per SM walks the
// (m, n) tile list; warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator,
// warps 4-7 and 8-11 = two epilogue groups.  The number of valid rows is read from device memory so the caller
// never synchronises on the packed token count.
#include <cudaTypedefs.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "ptx.cuh"

namespace ttr {

constexpr int GM = 128;            // tile rows  (UMMA M)
constexpr int GN = 128;            // tile cols  (UMMA N)
constexpr int GK = 32;             // fp32 elements per k-block = one 128-byte swizzle row
constexpr int G_MAX_STAGES = 8;
constexpr int G_A_BYTES = GM * GK * 4;   // 16 KB
constexpr int G_B_BYTES = GN * GK * 4;   // 16 KB
constexpr int G_STAGE_BYTES = G_A_BYTES + G_B_BYTES;
constexpr int G_ACC_COLS = GN;           // fp32 accumulator columns per buffer
constexpr int G_TMEM_COLS = 2 * G_ACC_COLS;
constexpr int G_THREADS = 384;             // warp 0 TMA, 1 MMA, 2 TMEM alloc, 4-7 and 8-11 two epilogue groups
constexpr int G_OUT_BYTES = GM * 32 * 4;   // staging buffer of one 128-row x 128-byte output chunk: 16 KB (x2 per group)
constexpr int G_F16_PITCH = GN * 2 + 16;   // F16: staging row of a whole 128-column fp16 tile + 16 B (conflict-free)
constexpr int G_F16_STAGE = GM * G_F16_PITCH;          // 34 KB per group
__host__ __device__ constexpr int stage_area(bool f16) { return f16 ? G_F16_STAGE : 2 * G_OUT_BYTES; }   // per epilogue group

extern int g_debug_flags;

// fp16 epilogue of one 128-row x 128-column block: TMEM -> registers (+bias from shared memory) -> row-major fp16
// staging.  The four tcgen05.ld of the block are double-buffered (chunk c + 1 is in flight while chunk c is converted)
// and the bias comes from a 512-byte shared copy of the tile's bias slice: r2 ncu of the K = 200 layer showed the
// epilogue groups, not loads / MMAs / DRAM, setting the tile rate (tensor pipe 28 %, DRAM 44 %, top stall
// long_scoreboard = the eight `__ldg(bias)` per 32-column chunk in front of every conversion).
// `release` is called right after the last tcgen05.ld of the block has completed.
template <typename Release>
__device__ __forceinline__ void f16_block_to_stage(uint32_t tmem_col0, int q, int r_in_tile, unsigned char* my_stage,
                                                   const float* __restrict__ bias_sm, bool last_block, Release release) {
  uint32_t ra[32], rb[32];
  const uint32_t lane_base = (uint32_t)(q * 32) << 16;
  ptx::tmem_ld_32x32(tmem_col0 + lane_base, ra);
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint32_t (&cur)[32] = (c & 1) ? rb : ra;
    uint32_t (&nxt)[32] = (c & 1) ? ra : rb;
    ptx::tmem_ld_wait();
    if (c < 3) ptx::tmem_ld_32x32(tmem_col0 + lane_base + 32 * (c + 1), nxt);
    else if (last_block) release();
    uint4* row = reinterpret_cast<uint4*>(my_stage + r_in_tile * G_F16_PITCH + c * 64);
#pragma unroll
    for (int j = 0; j < 4; ++j) {                   // four 16-byte chunks = 32 halves
      const float4 b0 = *reinterpret_cast<const float4*>(bias_sm + 32 * c + 8 * j);
      const float4 b1 = *reinterpret_cast<const float4*>(bias_sm + 32 * c + 8 * j + 4);
      const __half2 h0 = __floats2half2_rn(__uint_as_float(cur[8 * j + 0]) + b0.x, __uint_as_float(cur[8 * j + 1]) + b0.y);
      const __half2 h1 = __floats2half2_rn(__uint_as_float(cur[8 * j + 2]) + b0.z, __uint_as_float(cur[8 * j + 3]) + b0.w);
      const __half2 h2 = __floats2half2_rn(__uint_as_float(cur[8 * j + 4]) + b1.x, __uint_as_float(cur[8 * j + 5]) + b1.y);
      const __half2 h3 = __floats2half2_rn(__uint_as_float(cur[8 * j + 6]) + b1.z, __uint_as_float(cur[8 * j + 7]) + b1.w);
      row[j] = make_uint4(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1),
                          *reinterpret_cast<const uint32_t*>(&h2), *reinterpret_cast<const uint32_t*>(&h3));
    }
  }
}

// F16 = true: A, W and C are fp16 (kind::f16 MMAs, 64 halves per 128-byte k-block, fp32 accumulation and
// bias add, fp16 result) — the encode path's variant.  fp16 and tf32 carry the same 11-bit significand, so
// the products are as exact as the tf32 ones; what changes is the traffic: operands and the 6 KB/token gi
// row halve.  Measured with fp32 operands the kernel was bound by the SM's L2 read port on the operand
// tiles (524 KB per 128x128x512 tile = 59 B/clk/SM) and by the output stores, not by the tensor pipe.
//
// WRES = true (F16 only): "weight-stationary".  The grid is a multiple of the number of column tiles, so a
// CTA's column tile never changes; its W tile (all k-blocks, <= 128 KB at K = 512) is loaded once and stays in
// shared memory, and only A streams through the ring.  The operand loads were the limiter: one 128-byte row
// segment per ~2.5-4 cycles per SM (~50 B/clk), i.e. 4.5 k / 5.5 k cycles per 128x128 tile at K = 200 / 512
// while the MMAs need 1 k / 2 k — halving the streamed rows is worth more than a deeper ring.
template <bool F16, bool WRES>
__global__ void __launch_bounds__(G_THREADS, 1)
gemm_bias_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                      const __grid_constant__ CUtensorMap map_c, const float* __restrict__ bias, int m_bound,
                      const int32_t* __restrict__ m_valid, int N, int K, int dbg, void* __restrict__ c_out,
                      int n_stages, int n_groups) {
  extern __shared__ unsigned char smem_raw[];
  // [stages][A | B]; SWIZZLE_128B atoms need 1024-byte alignment in the shared window
  constexpr int GKE0 = F16 ? 2 * GK : GK;
  const int kb_count = ceil_div(K, GKE0);
  constexpr int STAGE_B = WRES ? G_A_BYTES : G_STAGE_BYTES;          // WRES: the ring holds A tiles only
  unsigned char* w_res = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);   // WRES: [k_blocks][16 KB]
  unsigned char* tiles = w_res + (WRES ? kb_count * G_B_BYTES : 0);
  unsigned char* out_stage = tiles + n_stages * STAGE_B;            // per epilogue group
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(out_stage + n_groups * stage_area(F16));
  uint64_t* empty_bar = full_bar + G_MAX_STAGES;
  uint64_t* acc_full = empty_bar + G_MAX_STAGES;    // [2] MMA -> epilogue
  uint64_t* acc_empty = acc_full + 2;           // [2] epilogue -> MMA
  uint64_t* w_full = acc_empty + 2;             // WRES: the resident W tile has landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int M = m_valid ? min(m_bound, *m_valid) : m_bound;
  const int m_tiles = ceil_div(M, GM), n_tiles = ceil_div(N, GN);
  const int total_tiles = m_tiles * n_tiles;
  constexpr int GKE = F16 ? 2 * GK : GK;            // elements per 128-byte k-block
  const int k_blocks = ceil_div(K, GKE);

  if (threadIdx.x == 0) {
    for (int s = 0; s < n_stages; ++s) { ptx::mbar_init(full_bar + s, 1); ptx::mbar_init(empty_bar + s, 1); }
    for (int b = 0; b < 2; ++b) { ptx::mbar_init(acc_full + b, 1); ptx::mbar_init(acc_empty + b, 4); }
    ptx::mbar_init(w_full, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_slot, G_TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer (whole warp runs the loop; one elected lane issues) =====
    if (ptx::elect_one()) {
      ptx::prefetch_tensormap(&map_a);
      ptx::prefetch_tensormap(&map_w);
    }
    if (WRES && blockIdx.x < total_tiles && ptx::elect_one()) {
      const int n0 = (blockIdx.x % n_tiles) * GN;         // constant for this CTA: gridDim.x % n_tiles == 0
      ptx::mbar_arrive_expect_tx(w_full, (uint32_t)(k_blocks * G_B_BYTES));
      for (int kb = 0; kb < k_blocks; ++kb) ptx::tma_load_2d(w_res + kb * G_B_BYTES, &map_w, kb * GKE, n0, w_full);
    }
    __syncwarp();
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int m0 = (tile / n_tiles) * GM, n0 = (tile % n_tiles) * GN;
      for (int kb = 0; kb < k_blocks; ++kb, ++it) {
        const int s = it % n_stages;
        const uint32_t ph = (uint32_t)(it / n_stages) & 1u;
        ptx::mbar_wait(empty_bar + s, ph ^ 1u);
        unsigned char* a_dst = tiles + s * STAGE_B;
        if (ptx::elect_one()) {
          ptx::mbar_arrive_expect_tx(full_bar + s, STAGE_B);
          ptx::tma_load_2d(a_dst, &map_a, kb * GKE, m0, full_bar + s);
          if (!WRES) ptx::tma_load_2d(a_dst + G_A_BYTES, &map_w, kb * GKE, n0, full_bar + s);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (whole warp runs the loop; one elected lane issues) =====
    constexpr uint32_t idesc = F16 ? ptx::make_idesc_f16(GM, GN) : ptx::make_idesc_tf32(GM, GN);
    int it = 0, local = 0;
    if (WRES && blockIdx.x < total_tiles) ptx::mbar_wait(w_full, 0u);
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++local) {
      const int buf = local & 1;
      const uint32_t aph = (uint32_t)(local >> 1) & 1u;
      ptx::mbar_wait(acc_empty + buf, aph ^ 1u);       // epilogue drained this accumulator
      ptx::tc_fence_after_sync();
      const uint32_t d_tmem = tmem_base + buf * G_ACC_COLS;
      for (int kb = 0; kb < k_blocks; ++kb, ++it) {
        const int s = it % n_stages;
        const uint32_t ph = (uint32_t)(it / n_stages) & 1u;
        ptx::mbar_wait(full_bar + s, ph);
        ptx::tc_fence_after_sync();
        const uint32_t a_addr = ptx::smem_u32(tiles + s * STAGE_B);
        const uint64_t a_desc = ptx::make_kmajor_sw128_desc(a_addr);
        const uint64_t b_desc = ptx::make_kmajor_sw128_desc(WRES ? ptx::smem_u32(w_res + kb * G_B_BYTES) : a_addr + G_A_BYTES);
        if (ptx::elect_one()) {
#pragma unroll
          for (int k = 0; k < GK / 8; ++k) {
            // k-steps that start at or beyond K multiply TMA zero fill only (K = 200: 13 of 16 fp16 steps are real)
            if ((kb * (GK / 8) + k) * (F16 ? 16 : 8) >= K) continue;
            // advance 8 tf32 / 16 halves = 32 bytes inside the 128-byte swizzle row: +2 in 16-byte units
            if (F16) ptx::mma_f16_ss(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
            else ptx::mma_tf32_ss(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
          }
          ptx::mma_commit(empty_bar + s);                 // smem slot free once these MMAs retire
          if (kb == k_blocks - 1) ptx::mma_commit(acc_full + buf);   // accumulator complete
        }
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue: TMEM -> registers (+bias) -> swizzled smem staging -> TMA store =====
    // (per-thread row stores hit 32 different 6 KB-strided rows per instruction and ran the
    // output at 1.1 TB/s, profiles/r1_encode_breakdown_v1.txt; the copy engine writes whole
    // 128-byte row segments.)  Rows between the valid count and m_bound are written too; they
    // belong to the caller's buffer and are never read.
    // Two groups of four warps alternate over the CTA's tiles (group g <-> accumulator buffer g): measured
    // with one group the ~4.5 k cycles of tcgen05.ld / bias / staging / store per tile, not the loads or the
    // MMAs, set the tile rate of the K = 200 layer.
    const int grp = (warp - 4) >> 2;
    const int q = (warp - 4) & 3;                         // TMEM lane quadrant of this warp
    const int r_in_tile = q * 32 + lane;
    unsigned char* my_stage = out_stage + grp * stage_area(F16);
    __shared__ float bias_tiles[2][GN];                   // per epilogue group: bias slice of the current column tile
    float* bias_sm = bias_tiles[grp];
    int bias_n0 = -1;
    int local = 0, chunk_no = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++local) {
      const int buf = local & 1;
      if (n_groups == 2 ? (buf != grp) : (grp != 0)) continue;
      const uint32_t aph = (uint32_t)(local >> 1) & 1u;
      const int m0 = (tile / n_tiles) * GM, n0 = (tile % n_tiles) * GN;
      ptx::mbar_wait(acc_full + buf, aph);
      ptx::tc_fence_after_sync();
      if constexpr (F16) {
        // fp16 result: the whole 128 x 128 tile is staged row-major (256 B + 16 B pad per row) and written with
        // plain coalesced stores, two full 256-byte row segments per warp instruction.  (TMA stores of
        // 128-byte-wide boxes — the SWIZZLE_128B limit — issue one 128 B segment per row and measured
        // 2.1 TB/s on this layout; per-thread row stores 1.1 TB/s.)
        if (n0 != bias_n0) {                              // (weight-stationary: once per CTA)
          ptx::named_bar_sync(1 + grp, 128);              // nobody still reads the previous slice
          const int n = n0 + r_in_tile;
          bias_sm[r_in_tile] = (bias && n < N) ? __ldg(bias + n) : 0.f;
          bias_n0 = n0;
          ptx::named_bar_sync(1 + grp, 128);
        }
        f16_block_to_stage(tmem_base + buf * G_ACC_COLS, q, r_in_tile, my_stage, bias_sm, true, [&]() {
          ptx::tc_fence_before_sync();                    // last read of this accumulator: hand it back
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(acc_empty + buf);
        });
        ptx::named_bar_sync(1 + grp, 128);
        if (!(dbg & 2048)) {
          __half* cbase = reinterpret_cast<__half*>(c_out);
          const int half_lane = lane & 15, sub = lane >> 4;
          const int ncol = n0 + 8 * half_lane;
#pragma unroll 4
          for (int rr = q * 32; rr < q * 32 + 32; rr += 2) {          // this warp's 32 rows, two per instruction
            const int rrow = rr + sub;
            const uint4 v = *reinterpret_cast<const uint4*>(my_stage + rrow * G_F16_PITCH + 16 * half_lane);
            if (m0 + rrow < M && ncol + 7 < N)
              *reinterpret_cast<uint4*>(cbase + (size_t)(m0 + rrow) * N + ncol) = v;
          }
        }
        ptx::named_bar_sync(1 + grp, 128);                // staging is free again for this group's next tile
        continue;
      }
      constexpr int CW = F16 ? 64 : 32;                   // columns per 128-byte staging row
#pragma unroll 1
      for (int c0 = 0; c0 < GN; c0 += CW, ++chunk_no) {
        uint32_t r[CW];
#pragma unroll
        for (int h = 0; h < CW / 32; ++h) {
          uint32_t (&rh)[32] = *reinterpret_cast<uint32_t(*)[32]>(r + 32 * h);
          ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * G_ACC_COLS + c0 + 32 * h, rh);
        }
        ptx::tmem_ld_wait();
        if (c0 + CW >= GN) {                              // last read of this accumulator: hand it back
          ptx::tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(acc_empty + buf);
        }
        unsigned char* stg = my_stage + (chunk_no & 1) * G_OUT_BYTES;
        // the store that used this staging buffer two chunks ago must have finished reading it
        if (q == 0 && lane == 0) ptx::bulk_wait_group_read<1>();
        ptx::named_bar_sync(1 + grp, 128);
        uint4* row = reinterpret_cast<uint4*>(stg + r_in_tile * 128);
        if (!(dbg & (1 << 18)))                                       // bit 18: (timing experiment) skip bias + staging
#pragma unroll
        for (int j = 0; j < 8; ++j) {                     // eight 16-byte chunks of the staging row
          constexpr int EPC = CW / 8;                     // elements per chunk: 4 floats or 8 halves
          const int n = n0 + c0 + EPC * j;
          float o[EPC];
#pragma unroll
          for (int e = 0; e < EPC; e += 4) {
            float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (bias && n + e + 3 < N) bv = __ldg(reinterpret_cast<const float4*>(bias + n + e));
            o[e + 0] = __uint_as_float(r[EPC * j + e + 0]) + bv.x;
            o[e + 1] = __uint_as_float(r[EPC * j + e + 1]) + bv.y;
            o[e + 2] = __uint_as_float(r[EPC * j + e + 2]) + bv.z;
            o[e + 3] = __uint_as_float(r[EPC * j + e + 3]) + bv.w;
          }
          uint4 pk;
          if (F16) {
            const __half2 h0 = __floats2half2_rn(o[0], o[1]), h1 = __floats2half2_rn(o[2], o[3]);
            const __half2 h2 = __floats2half2_rn(o[EPC - 4], o[EPC - 3]), h3 = __floats2half2_rn(o[EPC - 2], o[EPC - 1]);
            pk = make_uint4(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1),
                            *reinterpret_cast<const uint32_t*>(&h2), *reinterpret_cast<const uint32_t*>(&h3));
          } else {
            pk = make_uint4(__float_as_uint(o[0]), __float_as_uint(o[1]), __float_as_uint(o[2]), __float_as_uint(o[3]));
          }
          row[j ^ (r_in_tile & 7)] = pk;                  // SWIZZLE_128B: 16-byte chunk index XOR (row mod 8)
        }
        ptx::fence_proxy_async_smem();
        ptx::named_bar_sync(1 + grp, 128);
        if (q == 0 && lane == 0 && !(dbg & 2048)) {          // bit 11: timing experiment without the stores
          ptx::tma_store_2d(&map_c, stg, n0 + c0, m0);
          ptx::bulk_commit_group();
        }
      }
    }
    if (q == 0 && lane == 0) ptx::bulk_wait_group<0>();
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc(tmem_base, G_TMEM_COLS);
}

// ---- CTA-pair variant (fp16, K > 256): 256 x 256 output tiles on tcgen05.mma.cta_group::2 -----------------------
// The single-CTA kernel above streams a 128-row A block AND a 128-row W block per 128 x 128 x 64 step; at K = 512
// (no room for a resident W tile) that is 32 KB of operands per 2 MFLOP and the L2 -> SM fabric, not the tensor pipe,
// sets the rate (r1 ncu: 11.9 TB/s delivered to the SMs, tensor pipe 39 % active).  Here the two CTAs of a cluster
// compute ONE 256 x 256 tile: CTA r loads rows [m0 + 128 r, +128) of A and rows [n0 + 128 r, +128) of W — the same
// 32 KB per stage — and the pair's M = 256, N = 256 MMA reads both W halves, so each SM does twice the MMA work per
// operand byte it fetches.  Each CTA's tensor memory holds its 128 x 256 block of the tile (two buffers = all 512
// columns); its two epilogue groups write it out exactly like the single-CTA kernel, 128 columns at a time.
// Synchronisation as in the CTA-pair scorer (score_topk_mma.cu): plain TMA loads onto the CTA's own `full` barrier,
// the peer's warp 1 forwards each completion to the leader's `peer_full`, tcgen05.commit multicasts `stage free` and
// `accumulator full`, the peer's epilogue warps release accumulators with remote arrives on the leader's barrier.
constexpr int GP_ACC_COLS = 256;
constexpr int GP_TMEM_COLS = 512;
// The ring sets this kernel's speed: a stage is reloaded only after its MMAs retire, and a load takes ~3.5 k cycles
// under load, so stages / (3.5 k + 512) k-blocks complete per cycle (r2: 7,960 cycles per K = 512 tile with 4 stages,
// exactly that model).  The epilogue staging is therefore only 64 columns wide (18 KB per group instead of 34 KB):
// that buys a fifth 32 KB stage.
constexpr int GP_HALF = 64;                               // columns staged at a time
constexpr int GP_PITCH = GP_HALF * 2 + 16;                // staging row: 128 B + 16 B pad (conflict-free)
constexpr int GP_STAGE = GM * GP_PITCH;                   // 18 KB per epilogue group
constexpr int GP_STAGES = 5;

// 64 columns already in registers (two 32-column tcgen05.ld results) + bias -> row-major fp16 staging (GP_PITCH)
__device__ __forceinline__ void f16_regs_to_stage(const uint32_t (&ra)[32], const uint32_t (&rb)[32], int r_in_tile,
                                                  unsigned char* my_stage, const float* __restrict__ bias_sm) {
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    const uint32_t (&cur)[32] = c ? rb : ra;
    uint4* row = reinterpret_cast<uint4*>(my_stage + r_in_tile * GP_PITCH + c * 64);
#pragma unroll
    for (int j = 0; j < 4; ++j) {                   // four 16-byte chunks = 32 halves
      const float4 b0 = *reinterpret_cast<const float4*>(bias_sm + 32 * c + 8 * j);
      const float4 b1 = *reinterpret_cast<const float4*>(bias_sm + 32 * c + 8 * j + 4);
      const __half2 h0 = __floats2half2_rn(__uint_as_float(cur[8 * j + 0]) + b0.x, __uint_as_float(cur[8 * j + 1]) + b0.y);
      const __half2 h1 = __floats2half2_rn(__uint_as_float(cur[8 * j + 2]) + b0.z, __uint_as_float(cur[8 * j + 3]) + b0.w);
      const __half2 h2 = __floats2half2_rn(__uint_as_float(cur[8 * j + 4]) + b1.x, __uint_as_float(cur[8 * j + 5]) + b1.y);
      const __half2 h3 = __floats2half2_rn(__uint_as_float(cur[8 * j + 6]) + b1.z, __uint_as_float(cur[8 * j + 7]) + b1.w);
      row[j] = make_uint4(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1),
                          *reinterpret_cast<const uint32_t*>(&h2), *reinterpret_cast<const uint32_t*>(&h3));
    }
  }
}

// the same into a dense [128 rows][128 B] staging tile in the SWIZZLE_128B pattern of the output tensor map (16-byte
// chunk index XOR row mod 8): the source of a TMA store
__device__ __forceinline__ void f16_regs_to_swizzled(const uint32_t (&ra)[32], const uint32_t (&rb)[32], int r_in_tile,
                                                     unsigned char* my_stage, const float* __restrict__ bias_sm) {
  uint4* row = reinterpret_cast<uint4*>(my_stage + r_in_tile * 128);
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    const uint32_t (&cur)[32] = c ? rb : ra;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 b0 = *reinterpret_cast<const float4*>(bias_sm + 32 * c + 8 * j);
      const float4 b1 = *reinterpret_cast<const float4*>(bias_sm + 32 * c + 8 * j + 4);
      const __half2 h0 = __floats2half2_rn(__uint_as_float(cur[8 * j + 0]) + b0.x, __uint_as_float(cur[8 * j + 1]) + b0.y);
      const __half2 h1 = __floats2half2_rn(__uint_as_float(cur[8 * j + 2]) + b0.z, __uint_as_float(cur[8 * j + 3]) + b0.w);
      const __half2 h2 = __floats2half2_rn(__uint_as_float(cur[8 * j + 4]) + b1.x, __uint_as_float(cur[8 * j + 5]) + b1.y);
      const __half2 h3 = __floats2half2_rn(__uint_as_float(cur[8 * j + 6]) + b1.z, __uint_as_float(cur[8 * j + 7]) + b1.w);
      row[(4 * c + j) ^ (r_in_tile & 7)] = make_uint4(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1),
                                                      *reinterpret_cast<const uint32_t*>(&h2), *reinterpret_cast<const uint32_t*>(&h3));
    }
  }
}

// CL4 = true (with WRES; debug bit 16, NOT the default — measured slower, see the host code): clusters of FOUR CTAs =
// two pairs that share a 256-row block of A and own two adjacent
// 256-column tiles.  Each k-block of A is loaded ONCE per cluster row half — pair 0's CTA r issues the even k-blocks,
// pair 1's CTA r the odd ones — and multicast into both pairs' rings (`cp.async.bulk.tensor ... multicast::cluster`), so
// an SM requests half the A bytes per MMA: at K = 200 the A tile is re-read by six pairs and, together with the gi
// stores, the L2 <-> SM fabric was the limit (r2 probe: 0.42 ms with stores, 0.29 without, MMA floor 0.14 per 505 k tokens).
// A stage is refilled when BOTH pairs' MMAs have retired (`empty` counts two commits, each multicast to all four CTAs).
template <bool WRES, bool CL4>
__global__ void __launch_bounds__(G_THREADS, 1)
gemm_bias_pair_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                      const __grid_constant__ CUtensorMap map_c, const float* __restrict__ bias, int m_bound,
                      const int32_t* __restrict__ m_valid, int N, int K, int dbg, __half* __restrict__ c_out, int n_stages) {
  extern __shared__ unsigned char smem_raw[];
  constexpr int GKE = 2 * GK;                       // halves per 128-byte k-block
  constexpr int STAGE_B = WRES ? G_A_BYTES : G_STAGE_BYTES;
  const int k_blocks = ceil_div(K, GKE);
  unsigned char* w_res = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);   // WRES: [k_blocks][16 KB]
  unsigned char* tiles = w_res + (WRES ? k_blocks * G_B_BYTES : 0);
  unsigned char* out_stage = tiles + n_stages * STAGE_B;            // per epilogue group
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(out_stage + 2 * GP_STAGE);
  uint64_t* empty_bar = full_bar + G_MAX_STAGES;
  uint64_t* peer_full = empty_bar + G_MAX_STAGES;   // leader only: the peer's half of stage s has landed
  uint64_t* acc_full = peer_full + G_MAX_STAGES;    // [2] MMA -> epilogue
  uint64_t* acc_empty = acc_full + 2;               // [2] epilogue -> MMA (leader's: both CTAs' groups arrive)
  uint64_t* w_full = acc_empty + 2;                 // WRES: this CTA's resident W rows have landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 1);

  static_assert(!CL4 || WRES, "the A-multicast variant keeps W resident");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = ptx::cluster_ctarank();        // 0..1 (one pair) or 0..3 (two pairs)
  const uint32_t rank = crank & 1u;                     // rank inside the pair
  const uint32_t pr = CL4 ? (crank >> 1) : 0u;          // which pair of the cluster
  const uint32_t lead = crank & ~1u;                    // cluster rank of this pair's leader
  const bool leader = rank == 0u;
  const int M = m_valid ? min(m_bound, *m_valid) : m_bound;
  const int m_tiles = ceil_div(M, 2 * GM), n_tiles = ceil_div(N, 2 * GN);
  // work units: (256-row block, column tile) per pair, or (256-row block, two adjacent column tiles) per cluster of four
  const int n_units = CL4 ? n_tiles / 2 : n_tiles;
  const int total_tiles = m_tiles * n_units;
  const int pair = CL4 ? blockIdx.x >> 2 : blockIdx.x >> 1;
  const int n_pairs = CL4 ? gridDim.x >> 2 : gridDim.x >> 1;
  auto col_tile = [&](int tile) { return CL4 ? (tile % n_units) * 2 + (int)pr : tile % n_units; };
  auto row_tile = [&](int tile) { return tile / n_units; };

  if (threadIdx.x == 0) {
    for (int s = 0; s < n_stages; ++s) {
      ptx::mbar_init(full_bar + s, 1); ptx::mbar_init(empty_bar + s, CL4 ? 2 : 1); ptx::mbar_init(peer_full + s, 1);
    }
    for (int b = 0; b < 2; ++b) { ptx::mbar_init(acc_full + b, 1); ptx::mbar_init(acc_empty + b, 8); }
    ptx::mbar_init(w_full, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc_2cta(tmem_slot, GP_TMEM_COLS);
    ptx::tmem_relinquish_2cta();
  }
  ptx::tc_fence_before_sync();
  ptx::cluster_sync();                    // both CTAs' barriers exist before any multicast commit / remote arrive
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer: this CTA's 128 rows of A and 128 rows of W per k-block =====
    if (ptx::elect_one()) {
      ptx::prefetch_tensormap(&map_a);
      ptx::prefetch_tensormap(&map_w);
    }
    if (WRES && pair < total_tiles && ptx::elect_one()) {
      const int n0 = col_tile(pair) * 2 * GN + (int)rank * GN;        // constant for this pair: n_pairs % n_units == 0
      ptx::mbar_arrive_expect_tx(w_full, (uint32_t)(k_blocks * G_B_BYTES));
      for (int kb = 0; kb < k_blocks; ++kb) ptx::tma_load_2d(w_res + kb * G_B_BYTES, &map_w, kb * GKE, n0, w_full);
    }
    __syncwarp();
    int it = 0;
    for (int tile = pair; tile < total_tiles; tile += n_pairs) {
      const int m0 = row_tile(tile) * 2 * GM + (int)rank * GM, n0 = col_tile(tile) * 2 * GN + (int)rank * GN;
      for (int kb = 0; kb < k_blocks; ++kb, ++it) {
        const int s = it % n_stages;
        const uint32_t ph = (uint32_t)(it / n_stages) & 1u;
        ptx::mbar_wait(empty_bar + s, ph ^ 1u);
        unsigned char* a_dst = tiles + s * STAGE_B;
        if (ptx::elect_one()) {
          ptx::mbar_arrive_expect_tx(full_bar + s, STAGE_B);
          if (CL4) {
            // the two CTAs with the same rank-in-pair read the same A rows: they take turns loading a k-block for both
            // (data that lands in the other CTA before it has armed its barrier only drives its count negative)
            if ((uint32_t)(it & 1) == pr)
              ptx::tma_load_2d_multicast(a_dst, &map_a, kb * GKE, m0, full_bar + s, (uint16_t)((1u << rank) | (4u << rank)));
          } else {
            ptx::tma_load_2d(a_dst, &map_a, kb * GKE, m0, full_bar + s);
            if (!WRES) ptx::tma_load_2d(a_dst + G_A_BYTES, &map_w, kb * GKE, n0, full_bar + s);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1 && leader) {
    // ===== MMA issuer (leader CTA): M = 256 (128 rows per CTA), N = 256 (128 W rows per CTA), K = 16 =====
    constexpr uint32_t idesc = ptx::make_idesc_f16(2 * GM, 2 * GN);
    int it = 0, local = 0;
    // (the peer's W rows: its forwarder only reports stages after its own W has landed)
    if (WRES && pair < total_tiles) ptx::mbar_wait(w_full, 0u);
    for (int tile = pair; tile < total_tiles; tile += n_pairs, ++local) {
      const int buf = local & 1;
      const uint32_t aph = (uint32_t)(local >> 1) & 1u;
      ptx::mbar_wait(acc_empty + buf, aph ^ 1u);       // both CTAs' epilogues drained this accumulator
      ptx::tc_fence_after_sync();
      const uint32_t d_tmem = tmem_base + buf * GP_ACC_COLS;
      for (int kb = 0; kb < k_blocks; ++kb, ++it) {
        const int s = it % n_stages;
        const uint32_t ph = (uint32_t)(it / n_stages) & 1u;
        ptx::mbar_wait(full_bar + s, ph);
        ptx::mbar_wait(peer_full + s, ph);
        ptx::tc_fence_after_sync();
        const uint32_t a_addr = ptx::smem_u32(tiles + s * STAGE_B);
        const uint64_t a_desc = ptx::make_kmajor_sw128_desc(a_addr);
        const uint64_t b_desc = ptx::make_kmajor_sw128_desc(WRES ? ptx::smem_u32(w_res + kb * G_B_BYTES) : a_addr + G_A_BYTES);
        if (ptx::elect_one()) {
#pragma unroll
          for (int k = 0; k < GK / 8; ++k)
            if ((kb * (GK / 8) + k) * 16 < K)          // k-steps beyond K hold only zero fill (K = 200: 13 of 16)
              ptx::mma_f16_ss_2cta(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
          // stage s is free in both CTAs of the pair (CL4: one of the two commits all four CTAs wait for)
          ptx::mma_commit_2cta(empty_bar + s, CL4 ? (uint16_t)0xF : (uint16_t)3);
          if (kb == k_blocks - 1) ptx::mma_commit_2cta(acc_full + buf, (uint16_t)(3u << lead));
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ===== peer CTA: forward "my half of stage s has landed" to the leader's MMA issuer =====
    int it = 0;
    if (WRES && pair < total_tiles) ptx::mbar_wait(w_full, 0u);
    for (int tile = pair; tile < total_tiles; tile += n_pairs)
      for (int kb = 0; kb < k_blocks; ++kb, ++it) {
        const int s = it % n_stages;
        ptx::mbar_wait(full_bar + s, (uint32_t)(it / n_stages) & 1u);
        if (ptx::elect_one()) ptx::mbar_arrive_cluster(ptx::mapa(ptx::smem_u32(peer_full + s), lead));
        __syncwarp();
      }
  } else if (warp >= 4) {
    // ===== epilogue: this CTA's 128 x 256 block, four 64-column pieces through a staging tile =====
    const int grp = (warp - 4) >> 2;
    const int q = (warp - 4) & 3;                         // TMEM lane quadrant of this warp
    const int r_in_tile = q * 32 + lane;
    unsigned char* my_stage = out_stage + grp * GP_STAGE;
    __shared__ float bias_tiles[2][2 * GN];               // per epilogue group: bias slice of the current 256-column tile
    float* bias_sm = bias_tiles[grp];
    int bias_n0 = -1;
    int local = 0;
    for (int tile = pair; tile < total_tiles; tile += n_pairs, ++local) {
      const int buf = local & 1;
      if (buf != grp) continue;
      const uint32_t aph = (uint32_t)(local >> 1) & 1u;
      const int m0 = row_tile(tile) * 2 * GM + (int)rank * GM, nt0 = col_tile(tile) * 2 * GN;
      if (nt0 != bias_n0) {
        ptx::named_bar_sync(1 + grp, 128);                // nobody still reads the previous slice
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int n = nt0 + h * GN + r_in_tile;
          bias_sm[h * GN + r_in_tile] = (bias && n < N) ? __ldg(bias + n) : 0.f;
        }
        bias_n0 = nt0;
        ptx::named_bar_sync(1 + grp, 128);
      }
      ptx::mbar_wait(acc_full + buf, aph);
      ptx::tc_fence_after_sync();
      // Four 64-column pieces of the 256-column block, software-pipelined: the tcgen05.ld of piece c + 1 is in flight
      // while piece c is converted, staged and stored (TMEM reads run at 64 B/clk per SM: 512 cycles per piece that
      // used to sit in front of every conversion; r2 ncu of the K = 200 layer: epilogue groups, not MMAs or DRAM, set
      // the tile rate).  Debug bit 18: the un-pipelined loop (A/B).
      const uint32_t acc_col0 = tmem_base + buf * GP_ACC_COLS + ((uint32_t)(q * 32) << 16);
      auto release = [&]() {
        ptx::tc_fence_before_sync();                      // last read of this accumulator: hand it back
        __syncwarp();
        if (lane == 0) {
          if (leader) ptx::mbar_arrive(acc_empty + buf);
          else ptx::mbar_arrive_cluster(ptx::mapa(ptx::smem_u32(acc_empty + buf), lead));
        }
      };
      auto store_piece = [&](int n0) {
        ptx::named_bar_sync(1 + grp, 128);
        if (!(dbg & 2048)) {
          // 128-byte row segments: eight lanes per row, four rows per warp instruction
          const int chunk = lane & 7, sub = lane >> 3;
          const int ncol = n0 + 8 * chunk;
#pragma unroll 4
          for (int rr = q * 32; rr < q * 32 + 32; rr += 4) {
            const int rrow = rr + sub;
            const uint4 v = *reinterpret_cast<const uint4*>(my_stage + rrow * GP_PITCH + 16 * chunk);
            if (m0 + rrow < M && ncol + 7 < N)
              *reinterpret_cast<uint4*>(c_out + (size_t)(m0 + rrow) * N + ncol) = v;
          }
        }
        ptx::named_bar_sync(1 + grp, 128);                // staging is free again
      };
      uint32_t r0[2][32], r1[2][32];
      ptx::tmem_ld_32x32(acc_col0, r0[0]);
      ptx::tmem_ld_32x32(acc_col0 + 32, r1[0]);
#pragma unroll
      for (int ch = 0; ch < 2 * GN / GP_HALF; ++ch) {
        ptx::tmem_ld_wait();
        if (ch + 1 < 2 * GN / GP_HALF) {
          ptx::tmem_ld_32x32(acc_col0 + (ch + 1) * GP_HALF, r0[(ch + 1) & 1]);
          ptx::tmem_ld_32x32(acc_col0 + (ch + 1) * GP_HALF + 32, r1[(ch + 1) & 1]);
        } else {
          release();
        }
        if (dbg & (1 << 18)) {                              // debug bit 18 (A/B): row-major staging + coalesced stores
          f16_regs_to_stage(r0[ch & 1], r1[ch & 1], r_in_tile, my_stage, bias_sm + ch * GP_HALF);
          store_piece(nt0 + ch * GP_HALF);
        } else {
          // The piece leaves through ONE TMA store from a swizzled staging tile instead of 16 LDS + STG per thread;
          // the copy engine writes while the next piece is converted (r2 probe, 505 k tokens: K = 200 0.449 -> 0.419 ms,
          // K = 512 0.726 -> 0.671 ms).  Rows between the valid count and the end of the tile are written too
          // (caller-owned, never read), like the tf32 kernel's stores.
          if (q == 0 && lane == 0) ptx::bulk_wait_group_read<0>();      // the previous store has read the staging tile
          ptx::named_bar_sync(1 + grp, 128);
          f16_regs_to_swizzled(r0[ch & 1], r1[ch & 1], r_in_tile, my_stage, bias_sm + ch * GP_HALF);
          ptx::fence_proxy_async_smem();
          ptx::named_bar_sync(1 + grp, 128);
          if (q == 0 && lane == 0 && !(dbg & 2048)) {
            ptx::tma_store_2d(&map_c, my_stage, nt0 + ch * GP_HALF, m0);
            ptx::bulk_commit_group();
          }
        }
      }
    }
  }

  if (warp >= 4 && ((warp - 4) & 3) == 0 && lane == 0) ptx::bulk_wait_group<0>();    // (TMA-store variant) stores complete
  ptx::tc_fence_before_sync();
  ptx::cluster_sync();                    // the peer's MMAs / commits / remote arrives no longer target this CTA
  if (warp == 2) ptx::tmem_dealloc_2cta(tmem_base, GP_TMEM_COLS);
}

// ---- host: tensor maps ------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }
  return fn;
}

int make_rowmajor_map(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int box_rows, bool tf32);
int make_pitched_map(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t pitch, int box_rows,
                     bool tf32, bool atom32 = false);

// row-major fp32 [rows, cols] matrix, box = [box_rows, 32 floats], SWIZZLE_128B, OOB -> 0
int make_tf32_rowmajor_map(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int box_rows) {
  return make_rowmajor_map(map, base, rows, cols, box_rows, true);
}

int make_rowmajor_map(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int box_rows, bool tf32) {
  return make_pitched_map(map, base, rows, cols, cols, box_rows, tf32);
}

// [rows, cols] view of a row-major matrix with row pitch `pitch` floats (a column block of a wider matrix)
int make_pitched_map(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t pitch, int box_rows,
                     bool tf32, bool atom32) {
  auto enc = get_encode_fn();
  TTR_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)pitch * 4};
  cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  // TFLOAT32 lets the copy engine present tf32-typed data; bit 1 of the debug flags selects
  // plain FLOAT32 (hardware truncation in the MMA) for the rounding experiment in the tests.
  CUtensorMapDataType dt = (!tf32 || (g_debug_flags & 2)) ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_TFLOAT32;
  CUresult r = enc(map, dt, 2, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  TTR_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld cols=%lld)", (int)r,
              (long long)rows, (long long)cols);
  return TTR_OK;
}

// row-major fp16 [rows, cols] matrix, box = [box_rows, 64 halves], SWIZZLE_128B, OOB -> 0
int make_f16_rowmajor_map(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int box_rows) {
  auto enc = get_encode_fn();
  TTR_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  TTR_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (fp16) failed with CUresult %d (rows=%lld cols=%lld)", (int)r,
              (long long)rows, (long long)cols);
  return TTR_OK;
}

}  // namespace ttr

extern "C" int ttr_gemm_f16_bias(const void* A16, const void* W16, const float* bias, void* C16, int m_bound,
                                 const int32_t* m_valid, int N, int K, void* stream) {
  using namespace ttr;
  TTR_REQUIRE(m_bound >= 1 && N >= 1 && K >= 1, "ttr_gemm_f16_bias: bad shape");
  TTR_REQUIRE(K % 8 == 0 && N % 8 == 0, "ttr_gemm_f16_bias: K=%d and N=%d must be multiples of 8", K, N);
  TTR_REQUIRE(((uintptr_t)A16 & 15) == 0 && ((uintptr_t)W16 & 15) == 0 && ((uintptr_t)C16 & 15) == 0 &&
                  ((uintptr_t)bias & 15) == 0,
              "ttr_gemm_f16_bias: operands must be 16-byte aligned");
  CUtensorMap map_a, map_w, map_c;
  int rc = make_f16_rowmajor_map(&map_a, A16, m_bound, K, GM);
  if (rc != TTR_OK) return rc;
  rc = make_f16_rowmajor_map(&map_w, W16, N, K, GN);
  if (rc != TTR_OK) return rc;
  rc = make_f16_rowmajor_map(&map_c, C16, m_bound, N, GM);
  if (rc != TTR_OK) return rc;
  const int m_tiles = ceil_div(m_bound, GM), n_tiles = ceil_div(N, GN);
  const int tiles = m_tiles * n_tiles;
  const int k_blocks = ceil_div(K, 2 * GK);
  const size_t misc = (2 * G_MAX_STAGES + 5) * sizeof(uint64_t) + 16 + 1024;
  if (!(g_debug_flags & (1 << 30)) && sm_count() >= 2) {
    // Default for every K: CTA pairs, 256 x 256 tiles (layer 0, K = 200: 636 vs 523 TFLOP/s for the weight-stationary
    // single-CTA kernel; layer 1, K = 512: 1,160 vs 727).  Debug bit 30 selects the single-CTA kernels below (A/B).
    const size_t misc_p = (3 * G_MAX_STAGES + 5) * sizeof(uint64_t) + 16 + 1024;
    const size_t smem_max = 227 * 1024 - 2048;          // the kernel also has 2 KB of static shared memory (bias slices)
    const int pn_tiles = ceil_div(N, 2 * GN), pm_tiles = ceil_div(m_bound, 2 * GM);
    const int pair_tiles = pm_tiles * pn_tiles;
    // weight-stationary pairs when a CTA's 128 W rows (all k-blocks) fit next to >= 6 A stages and every column tile
    // gets at least one pair (K <= 256); debug bit 19: stream W through the ring as for K = 512 (A/B)
    int res_stages = 0;
    for (int st = G_MAX_STAGES; st >= 6 && !res_stages; --st)
      if ((size_t)k_blocks * G_B_BYTES + (size_t)st * G_A_BYTES + 2 * (size_t)GP_STAGE + misc_p <= smem_max) res_stages = st;
    const bool wres = res_stages > 0 && pn_tiles <= sm_count() / 2 && !(g_debug_flags & (1 << 19));
    const int STP = wres ? res_stages : GP_STAGES;
    const size_t smem_p = wres ? (size_t)k_blocks * G_B_BYTES + (size_t)STP * G_A_BYTES + 2 * (size_t)GP_STAGE + misc_p
                               : (size_t)STP * G_STAGE_BYTES + 2 * (size_t)GP_STAGE + misc_p;
    // debug bit 16 (A/B, weight-stationary shapes with an even number of column tiles): two pairs per cluster sharing
    // (multicasting) the A block.  Measured SLOWER — 0.707 vs 0.421 ms per 505 k tokens at K = 200 (0.577 vs 0.286
    // without the stores): halving the A requests does not pay for refilling a stage only after BOTH pairs have retired
    // it and for the multicast's latency — so independent pairs stay the default.
    const bool cl4 = wres && pn_tiles % 2 == 0 && pn_tiles / 2 <= sm_count() / 4 && (g_debug_flags & (1 << 16));
    const void* kern = cl4 ? (const void*)gemm_bias_pair_kernel<true, true>
                           : wres ? (const void*)gemm_bias_pair_kernel<true, false> : (const void*)gemm_bias_pair_kernel<false, false>;
    static thread_local int attr_dev = -1;
    int cur_dev = 0;
    TTR_CHECK_CUDA(cudaGetDevice(&cur_dev));
    if (attr_dev != cur_dev) {
      TTR_CHECK_CUDA(cudaFuncSetAttribute(gemm_bias_pair_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
      TTR_CHECK_CUDA(cudaFuncSetAttribute(gemm_bias_pair_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
      TTR_CHECK_CUDA(cudaFuncSetAttribute(gemm_bias_pair_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
      attr_dev = cur_dev;
    }
    int pairs = std::max(1, std::min(pair_tiles, sm_count() / 2));
    if (wres) pairs = pn_tiles * std::max(1, std::min(sm_count() / 2 / pn_tiles, pm_tiles));   // multiple of the column tiles
    const int cl_size = cl4 ? 4 : 2;
    if (cl4) {                                            // clusters of four: a multiple of the column-tile PAIRS
      const int groups = pn_tiles / 2;
      pairs = 2 * groups * std::max(1, std::min(sm_count() / 4 / groups, pm_tiles));
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * pairs);
    cfg.blockDim = dim3(G_THREADS);
    cfg.dynamicSmemBytes = smem_p;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cl_size; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    int dbgv = g_debug_flags, stp = STP;
    __half* c16 = reinterpret_cast<__half*>(C16);
    void* args[] = {(void*)&map_a, (void*)&map_w, (void*)&map_c, (void*)&bias, (void*)&m_bound, (void*)&m_valid, (void*)&N,
                    (void*)&K, (void*)&dbgv, (void*)&c16, (void*)&stp};
    TTR_CHECK_CUDA(cudaLaunchKernelExC(&cfg, kern, args));
    return TTR_OK;
  }
  // weight-stationary when the W tile fits next to a >= 5-deep A ring and two epilogue groups (K <= 256) and every
  // column tile gets at least one CTA; at K = 512 only a 3-deep ring would fit and that measured slower than
  // streaming both operands through a 5-deep ring (2.29 vs 1.80 ms per 1.0 M tokens)
  int ws_stages = 0;
  for (int st = G_MAX_STAGES; st >= 5 && !ws_stages; --st)
    if ((size_t)k_blocks * G_B_BYTES + (size_t)st * G_A_BYTES + 2 * stage_area(true) + misc <= 227 * 1024) ws_stages = st;
  if (ws_stages && n_tiles <= sm_count() && !(g_debug_flags & (1 << 19))) {
    const size_t smem_res = (size_t)k_blocks * G_B_BYTES + (size_t)ws_stages * G_A_BYTES + 2 * stage_area(true) + misc;
    const int per_n = std::max(1, std::min(sm_count() / n_tiles, m_tiles));
    TTR_CHECK_CUDA(cudaFuncSetAttribute(gemm_bias_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_res));
    gemm_bias_kernel<true, true><<<n_tiles * per_n, G_THREADS, smem_res, (cudaStream_t)stream>>>(
        map_a, map_w, map_c, bias, m_bound, m_valid, N, K, g_debug_flags, C16, ws_stages, 2);
    TTR_CHECK_LAUNCH();
    return TTR_OK;
  }
  constexpr int ST16 = 4;                  // 4 x 32 KB ring + two 34 KB staging groups (1.80 ms per 1.0 M tokens at
                                           // K = 512; 5 stages + one group measured 2.00 ms)
  const size_t smem = (size_t)ST16 * G_STAGE_BYTES + 2 * stage_area(true) + misc;
  TTR_CHECK_CUDA(cudaFuncSetAttribute(gemm_bias_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = std::min(tiles, sm_count());
  gemm_bias_kernel<true, false><<<grid, G_THREADS, smem, (cudaStream_t)stream>>>(map_a, map_w, map_c, bias, m_bound, m_valid, N, K,
                                                                                g_debug_flags, C16, ST16, 2);
  TTR_CHECK_LAUNCH();
  return TTR_OK;
}

extern "C" int ttr_gemm_tf32_bias(const float* A, const float* W, const float* bias, float* C, int m_bound,
                                  const int32_t* m_valid, int N, int K, void* stream) {
  using namespace ttr;
  TTR_REQUIRE(m_bound >= 1 && N >= 1 && K >= 1, "ttr_gemm_tf32_bias: bad shape");
  TTR_REQUIRE(K % 4 == 0 && N % 4 == 0, "ttr_gemm_tf32_bias: K=%d and N=%d must be multiples of 4", K, N);
  TTR_REQUIRE(((uintptr_t)A & 15) == 0 && ((uintptr_t)W & 15) == 0 && ((uintptr_t)C & 15) == 0 &&
                  ((uintptr_t)bias & 15) == 0,
              "ttr_gemm_tf32_bias: operands must be 16-byte aligned");
  CUtensorMap map_a, map_w, map_c;
  int rc = make_tf32_rowmajor_map(&map_a, A, m_bound, K, GM);
  if (rc != TTR_OK) return rc;
  rc = make_tf32_rowmajor_map(&map_w, W, N, K, GN);
  if (rc != TTR_OK) return rc;
  rc = make_rowmajor_map(&map_c, C, m_bound, N, GM, false);
  if (rc != TTR_OK) return rc;
  constexpr int ST32 = 6;                  // 6 x 32 KB ring + one 32 KB staging group (as measured best for tf32)
  const size_t smem = (size_t)ST32 * G_STAGE_BYTES + stage_area(false) + (2 * G_MAX_STAGES + 5) * sizeof(uint64_t) + 16 + 1024;
  TTR_CHECK_CUDA(cudaFuncSetAttribute(gemm_bias_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int tiles = ceil_div(m_bound, GM) * ceil_div(N, GN);
  const int grid = std::min(tiles, sm_count());
  gemm_bias_kernel<false, false><<<grid, G_THREADS, smem, (cudaStream_t)stream>>>(map_a, map_w, map_c, bias, m_bound, m_valid, N, K,
                                                                          g_debug_flags, C, ST32, 1);
  TTR_CHECK_LAUNCH();
  return TTR_OK;
}
