// score_topk_mma.cu — exact cosine top-k for query batches > 8: tcgen05 score GEMM fused
// with a per-query streaming top-k, never materialising [B, N].
//
// Replaces `torch.matmul(q, D.t())` + `torch.topk` (backend/evaluators.py:185-186) for the
// batched configs (B = 256, 4096).  Per CTA:
//   * 128 queries (UMMA M = 128), rounded to tf32, live in TENSOR MEMORY as the A operand for
//     the whole kernel (256 columns), which leaves shared memory to the document stream;
//   * 32-document tiles (32 KB) stream through a 3-stage TMA ring (SWIZZLE_128B, TFLOAT32 maps:
//     the copy engine rounds operands to nearest-even); one elected thread issues kind::tf32
//     MMAs (A from TMEM, B from shared memory) into one of four 32-column TMEM accumulators;
//   * four epilogue warps: thread = query.  One tcgen05.ld brings the thread its 32 scores;
//     they are filtered against max(own k-th best, a global lower bound on the k-th best that
//     all CTAs share through an atomic max) and survivors are appended to the thread's
//     private 128-slot list in shared memory ([slot][thread] layout, conflict-free).  When a
//     list could overflow, all 32 lanes of the warp sort their own lists at once with a
//     data-independent bitonic network (SIMT-parallel, no shuffles) and keep the best k.
// HBM traffic: N * 1024 B per 128-query pass (document tiles are shared across query tiles
// through L2 when B > 128).  v1 of this kernel (Q in shared memory, lists in L2, warp-
// cooperative compaction) spent ~90% of its issue slots in the compaction; see
// profiles/r1_score_topk_mma_v1_ncu_raw.csv.
#include <cudaTypedefs.h>

#include "common.cuh"
#include "ptx.cuh"
#include "topk_common.cuh"

namespace ttr {

int make_tf32_rowmajor_map(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int box_rows);

constexpr int SM_DIM = 256;
constexpr int SM_MQ = 128;                 // queries per CTA (UMMA M)
constexpr int SM_ND = 32;                  // documents per tile (UMMA N)
constexpr int SM_KB = 8;                   // k-blocks of 32 floats
constexpr int SM_STAGES = 3;
constexpr int SM_D_KB_BYTES = SM_ND * 128;              // one k-block of a doc tile: 4 KB
constexpr int SM_STAGE_BYTES = SM_ND * SM_DIM * 4;      // 32 KB
constexpr int SM_NACC = 2;                              // TMEM accumulator buffers (tiles in flight)
constexpr int SM_KSPLIT = 4;                            // independent accumulation chains per tile
constexpr int SM_ACC_COLS = SM_KSPLIT * SM_ND;          // columns per buffer: 4 partial sums x 32 docs
constexpr int SM_Q_COLS = SM_DIM;                       // A operand: one column per k element
constexpr int SM_TMEM_COLS = 512;
constexpr int SM_THREADS = 256;
constexpr int SM_CAP = TOPK_CAP;                        // candidate slots per query
constexpr int SM_LIST_BYTES = SM_CAP * SM_MQ * 4;       // one array ([slot][thread]): 64 KB

__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
  if (v >= 0.f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

// Every lane sorts ITS OWN 128-slot list (descending by (score, idx)); lists are interleaved
// [slot][thread] so a warp's accesses to one slot are 32 consecutive words.
__device__ __forceinline__ void thread_sort128_desc(float* ls, int32_t* li) {
#pragma unroll 1
  for (int k = 2; k <= SM_CAP; k <<= 1) {
#pragma unroll 1
    for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll 4
      for (int t = 0; t < SM_CAP / 2; ++t) {
        const int a = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int b = a | j;
        const bool desc = ((a & k) == 0);
        const float sa = ls[a * SM_MQ], sb = ls[b * SM_MQ];
        const int32_t ia = li[a * SM_MQ], ib = li[b * SM_MQ];
        const bool b_better = key_better<int32_t>(sb, ib, sa, ia);
        if (b_better == desc) {
          ls[a * SM_MQ] = sb; ls[b * SM_MQ] = sa;
          li[a * SM_MQ] = ib; li[b * SM_MQ] = ia;
        }
      }
    }
  }
}

// Optional timeline trace of CTA (0,0): trace[role][tile] = clock64 at a pipeline event
// (roles: 0 producer issued, 1 MMA saw full, 2 MMA issued+committed, 3 epilogue saw acc_full,
// 4 epilogue released the accumulator).  Test/diagnostic only (ttr_debug_set_trace).
long long* g_score_trace = nullptr;
constexpr int SM_TRACE_TILES = 256;
#define SM_TRACE(role, it)                                                                      \
  do {                                                                                          \
    if (trace && blockIdx.x == 0 && blockIdx.y == 0 && (it) < SM_TRACE_TILES && lane == 0)      \
      trace[(role) * SM_TRACE_TILES + (it)] = clock64();                                        \
  } while (0)

__global__ void init_tau_kernel(float* tau, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) tau[i] = -INFINITY;
}

__global__ void __launch_bounds__(SM_THREADS, 1)
score_topk_mma_kernel(const float* __restrict__ Q, const __grid_constant__ CUtensorMap map_d, int B, int64_t N,
                      int k, int n_slices, float* __restrict__ tau_g, float* __restrict__ part_s,
                      int32_t* __restrict__ part_i, long long* __restrict__ trace) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* base = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* ring = base;                                      // [stages][SM_KB][32 rows][128 B]
  float* list_s = reinterpret_cast<float*>(ring + SM_STAGES * SM_STAGE_BYTES);   // [SM_CAP][128]
  int32_t* list_i = reinterpret_cast<int32_t*>(reinterpret_cast<unsigned char*>(list_s) + SM_LIST_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(list_i) + SM_LIST_BYTES);
  uint64_t* empty_bar = full_bar + SM_STAGES;
  uint64_t* acc_full = empty_bar + SM_STAGES;
  uint64_t* acc_empty = acc_full + SM_NACC;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + SM_NACC);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x;                 // query tile
  const int slice = blockIdx.y;              // document slice
  const int q0 = qt * SM_MQ;
  const int64_t n_tiles = ceil_div64(N, (int64_t)SM_ND);

  if (threadIdx.x == 0) {
    for (int s = 0; s < SM_STAGES; ++s) { ptx::mbar_init(full_bar + s, 1); ptx::mbar_init(empty_bar + s, 1); }
    for (int b = 0; b < SM_NACC; ++b) { ptx::mbar_init(acc_full + b, 1); ptx::mbar_init(acc_empty + b, 4); }
    ptx::fence_mbar_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_slot, SM_TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_acc = tmem_base + SM_Q_COLS;

  // ---- stage the query tile into tensor memory: thread (warp 4+w, lane) owns TMEM lane 32w+lane
  if (warp >= 4) {
    const int qw = warp - 4;
    const int q = q0 + qw * 32 + lane;
    const float4* src = reinterpret_cast<const float4*>(Q + (int64_t)(q < B ? q : 0) * SM_DIM);
#pragma unroll 1
    for (int c0 = 0; c0 < SM_DIM; c0 += 32) {
      uint32_t r[32];
#pragma unroll
      for (int v = 0; v < 8; ++v) {
        float4 x = (q < B) ? __ldg(src + (c0 >> 2) + v) : make_float4(0.f, 0.f, 0.f, 0.f);
        r[4 * v + 0] = __float_as_uint(round_tf32(x.x));
        r[4 * v + 1] = __float_as_uint(round_tf32(x.y));
        r[4 * v + 2] = __float_as_uint(round_tf32(x.z));
        r[4 * v + 3] = __float_as_uint(round_tf32(x.w));
      }
      ptx::tmem_st_32x32(tmem_base + ((uint32_t)(qw * 32) << 16) + c0, r);
    }
    ptx::tmem_st_wait();
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();

  if (warp == 0) {
    // ===== TMA producer (whole warp runs the loop; one elected lane issues) =====
    if (ptx::elect_one()) ptx::prefetch_tensormap(&map_d);
    int it = 0;
    for (int64_t t = slice; t < n_tiles; t += n_slices, ++it) {
      const int s = it % SM_STAGES;
      const uint32_t ph = (uint32_t)(it / SM_STAGES) & 1u;
      ptx::mbar_wait(empty_bar + s, ph ^ 1u);
      unsigned char* dst = ring + s * SM_STAGE_BYTES;
      const int32_t d0 = (int32_t)(t * SM_ND);
      if (ptx::elect_one()) {
        ptx::mbar_arrive_expect_tx(full_bar + s, SM_STAGE_BYTES);
#pragma unroll
        for (int kb = 0; kb < SM_KB; ++kb)
          ptx::tma_load_2d(dst + kb * SM_D_KB_BYTES, &map_d, kb * 32, d0, full_bar + s);
      }
      __syncwarp();
      SM_TRACE(0, it);
    }
  } else if (warp == 1) {
    // ===== MMA issuer (whole warp runs the loop; one elected lane issues) =====
    constexpr uint32_t idesc = ptx::make_idesc_tf32(SM_MQ, SM_ND);
    int it = 0;
    for (int64_t t = slice; t < n_tiles; t += n_slices, ++it) {
      const int s = it % SM_STAGES;
      const uint32_t ph = (uint32_t)(it / SM_STAGES) & 1u;
      const int buf = it % SM_NACC;
      const uint32_t aph = (uint32_t)(it / SM_NACC) & 1u;
      ptx::mbar_wait(acc_empty + buf, aph ^ 1u);
      ptx::mbar_wait(full_bar + s, ph);
      ptx::tc_fence_after_sync();
      SM_TRACE(1, it);
      const uint32_t d_addr = ptx::smem_u32(ring + s * SM_STAGE_BYTES);
      const uint32_t d_tmem = tmem_acc + buf * SM_ACC_COLS;
      const uint64_t b_desc0 = ptx::make_kmajor_sw128_desc(d_addr);
      if (ptx::elect_one()) {
        // The K = 256 reduction is split into SM_KSPLIT independent accumulation chains
        // (k-steps [c*8, c*8+8) -> partial accumulator c) issued round-robin so consecutive
        // MMAs never depend on each other; the epilogue adds the partial sums.
#pragma unroll
        for (int kk = 0; kk < 32 / SM_KSPLIT; ++kk) {
#pragma unroll
          for (int c = 0; c < SM_KSPLIT; ++c) {
            const int ks = c * (32 / SM_KSPLIT) + kk;              // k-step 0..31 (8 floats each)
            // k-block (ks>>2) starts (ks>>2)*4 KB further (>>4 in descriptor units); 32 B per k-step inside it
            const uint64_t b_desc = b_desc0 + (uint64_t)((ks >> 2) * (SM_D_KB_BYTES >> 4) + 2 * (ks & 3));
            ptx::mma_tf32_ts(d_tmem + c * SM_ND, tmem_base + ks * 8, b_desc, idesc, kk != 0);
          }
        }
        ptx::mma_commit(empty_bar + s);
        ptx::mma_commit(acc_full + buf);
      }
      __syncwarp();
      SM_TRACE(2, it);
    }
  } else if (warp >= 4) {
    // ===== epilogue: one thread per query =====
    const int qw = warp - 4;
    const int ql = qw * 32 + lane;             // query inside the tile == TMEM lane == list column
    const int q = q0 + ql;
    const bool q_valid = q < B;
    float* ls = list_s + ql;
    int32_t* li = list_i + ql;
    float tau_s = -INFINITY;
    int32_t tau_i = IDX_PAD;
    int cnt = 0;
    int it = 0;
    // Global lower bound on the k-th best.  An L2 read under a saturated memory system costs
    // microseconds, so it is refreshed every 8 tiles and consumed one refresh later (the load
    // is in flight for 8 tiles and never sits on the per-tile critical path).
    float tg = q_valid ? -INFINITY : INFINITY;
    float tg_pending = tg;
    for (int64_t t = slice; t < n_tiles; t += n_slices, ++it) {
      const int buf = it % SM_NACC;
      const uint32_t aph = (uint32_t)(it / SM_NACC) & 1u;
      if ((it & 7) == 0 && q_valid) {
        tg = fmaxf(tg, tg_pending);
        tg_pending = __ldcg(tau_g + q);
      }
      ptx::mbar_wait(acc_full + buf, aph);
      ptx::tc_fence_after_sync();
      if (qw == 0) SM_TRACE(3, it);
      float sc32[32];
      {
        uint32_t r[32];
        const uint32_t tcol = tmem_acc + ((uint32_t)(qw * 32) << 16) + buf * SM_ACC_COLS;
        ptx::tmem_ld_32x32(tcol, r);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) sc32[j] = __uint_as_float(r[j]);
#pragma unroll
        for (int c = 1; c < SM_KSPLIT; ++c) {
          ptx::tmem_ld_32x32(tcol + c * SM_ND, r);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) sc32[j] += __uint_as_float(r[j]);
        }
      }
      ptx::tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(acc_empty + buf);
      if (qw == 0) SM_TRACE(4, it);
      const int64_t d0 = t * SM_ND;
      const float thr = fmaxf(tg, tau_s);
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float sc = sc32[j];
        if (sc >= thr) {
          const int64_t doc = d0 + j;
          if (doc < N && key_better<int32_t>(sc, (int32_t)doc, tau_s, tau_i)) {
            ls[cnt * SM_MQ] = sc;
            li[cnt * SM_MQ] = (int32_t)doc;
            ++cnt;
          }
        }
      }
      if (__any_sync(0xffffffffu, cnt > SM_CAP - SM_ND)) {
        // all 32 lanes compact their own list together (uniform control flow)
        for (int e = cnt; e < SM_CAP; ++e) { ls[e * SM_MQ] = -INFINITY; li[e * SM_MQ] = IDX_PAD; }
        thread_sort128_desc(ls, li);
        if (cnt >= k) {
          cnt = k;
          tau_s = ls[(k - 1) * SM_MQ];
          tau_i = li[(k - 1) * SM_MQ];
          if (q_valid) atomic_max_float(tau_g + q, tau_s);
        }
      }
    }
    // final sort, then emit this CTA's partial list for the query
    for (int e = cnt; e < SM_CAP; ++e) { ls[e * SM_MQ] = -INFINITY; li[e * SM_MQ] = IDX_PAD; }
    thread_sort128_desc(ls, li);
    if (q_valid) {
      const int64_t pbase = ((int64_t)q * n_slices + slice) * k;
      for (int e = 0; e < k; ++e) {
        part_s[pbase + e] = ls[e * SM_MQ];
        part_i[pbase + e] = li[e * SM_MQ];
      }
    }
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc(tmem_base, SM_TMEM_COLS);
}

struct MmaPlan {
  int n_qt, n_slices;
  int64_t tau_off, part_s_off, part_i_off, total;
};

MmaPlan mma_plan(int B, int k) {
  MmaPlan p;
  p.n_qt = ceil_div(B, SM_MQ);
  const int sms = sm_count();
  p.n_slices = std::max(1, sms / p.n_qt);      // one wave: n_qt * n_slices <= #SMs (1 CTA per SM)
  const int64_t bp = (int64_t)p.n_qt * SM_MQ;
  p.tau_off = 0;
  p.part_s_off = (bp * 4 + 255) / 256 * 256;
  const int64_t part_elems = bp * p.n_slices * k;
  p.part_i_off = p.part_s_off + part_elems * 4;
  p.total = p.part_i_off + part_elems * 4 + 256;
  return p;
}

int launch_score_topk_mma(const float* Q, int B, const float* docs, int64_t N, int k, void* workspace,
                          float** part_s_out, int32_t** part_i_out, int* parts_out, cudaStream_t st) {
  MmaPlan p = mma_plan(B, k);
  unsigned char* ws = reinterpret_cast<unsigned char*>(workspace);
  float* tau = reinterpret_cast<float*>(ws + p.tau_off);
  float* part_s = reinterpret_cast<float*>(ws + p.part_s_off);
  int32_t* part_i = reinterpret_cast<int32_t*>(ws + p.part_i_off);
  CUtensorMap map_d;
  int rc = make_tf32_rowmajor_map(&map_d, docs, N, SM_DIM, SM_ND);
  if (rc != TTR_OK) return rc;
  const int nq_pad = p.n_qt * SM_MQ;
  init_tau_kernel<<<ceil_div(nq_pad, 256), 256, 0, st>>>(tau, nq_pad);
  TTR_CHECK_LAUNCH();
  const size_t smem = (size_t)SM_STAGES * SM_STAGE_BYTES + 2 * SM_LIST_BYTES + (2 * SM_STAGES + 2 * SM_NACC) * 8 + 16 + 1024;
  TTR_CHECK_CUDA(cudaFuncSetAttribute(score_topk_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(p.n_qt, p.n_slices);
  score_topk_mma_kernel<<<grid, SM_THREADS, smem, st>>>(Q, map_d, B, N, k, p.n_slices, tau, part_s, part_i,
                                                        g_score_trace);
  TTR_CHECK_LAUNCH();
  *part_s_out = part_s;
  *part_i_out = part_i;
  *parts_out = p.n_slices;
  return TTR_OK;
}

int64_t score_topk_mma_workspace_bytes(int B, int k) { return mma_plan(B, k).total; }

}  // namespace ttr

// diagnostic: device buffer of 5 * 256 int64 that receives the pipeline timeline of CTA (0,0), or NULL
extern "C" int ttr_debug_set_trace(long long* trace) {
  ttr::g_score_trace = trace;
  return TTR_OK;
}
