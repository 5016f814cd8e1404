// score_topk_mma.cu — exact cosine top-k for query batches > 4: tcgen05 score GEMM fused
// with a per-query streaming top-k, never materialising [B, N].
//
// Replaces `torch.matmul(q, D.t())` + `torch.topk` (backend/evaluators.py:185-186) for the
// batched configs (B = 256, 4096).  Per CTA:
//   * 128 queries (UMMA M = 128), rounded to tf32, live in TENSOR MEMORY as the A operand for
//     the whole kernel (256 columns), which leaves shared memory to the document stream;
//   * 32-document tiles (32 KB) stream through a 5-stage TMA ring (160 KB in flight per SM:
//     the measured TMA latency under full HBM load is ~4800 cycles, profiles/r1_*trace*);
//     TFLOAT32 tensor maps make the copy engine round operands to nearest-even;
//   * one elected thread issues kind::tf32 MMAs (A from TMEM, B from shared memory, N = 32,
//     ~23 cycles each, profiles/r1_mma_probe.txt) into one of four TMEM accumulators;
//   * four epilogue warps, thread = query: one tcgen05.ld brings the thread its 32 scores;
//     they are filtered against max(own k-th-best bound, a global bound shared by all CTAs
//     through an atomic max); survivors are appended to the thread's private 128-slot list
//     (scores in shared memory, [slot][thread] layout; document ids in an L2-resident scratch
//     with the same layout).  When a list could overflow, the lane finds a pivot with
//     k <= #(score >= pivot) <= 64 by bisection over the float key space (count passes only:
//     independent shared-memory loads, no sorting) and compacts in place.
// The kernel emits unsorted lists of <= 64 candidates per (CTA, query); the merge kernel
// produces the sorted global top-k.  HBM traffic: N * 1024 B per 128-query pass (document
// tiles are shared across query tiles through L2 when B > 128).
//
// History (profiles/): v1 kept Q in shared memory and lists in L2 with warp-cooperative
// sorts (90 % of issue slots in the sort); v2/v3 sorted per thread in shared memory (154 k
// cycles per round, 3 stages: 1850 cycles per tile); this is v4.
#include <cudaTypedefs.h>

#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "ptx.cuh"
#include "topk_common.cuh"

namespace ttr {

int make_tf32_rowmajor_map(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int box_rows);
extern int g_debug_flags;   // bits 3-5: timing experiments that break results (skip LDTM / pin the B operand)

constexpr int SM_DIM = 256;
constexpr int SM_MQ = 128;                 // queries per CTA (UMMA M)
constexpr int SM_ND = 32;                  // documents per tile (UMMA N)
constexpr int SM_KB = 8;                   // k-blocks of 32 floats
constexpr int SM_STAGES = 4;
constexpr int SM_D_KB_BYTES = SM_ND * 128;              // one k-block of a doc tile: 4 KB
constexpr int SM_STAGE_BYTES = SM_ND * SM_DIM * 4;      // 32 KB
constexpr int SM_NACC_MAX = 8;                          // TMEM accumulator buffers (tiles in flight): 8 x 32 columns next to the
                                                        // 256 query columns for single CTAs, 4 x 64 for CTA pairs.  With 4 a burst of
                                                        // tiles with survivors (a keeper needs ~1,400 cycles for one, a tile arrives
                                                        // every ~1,000) stalled the MMA issuer and through it the TMA ring: 1,320 cycles
                                                        // per tile on a 1.1 M-document shard, where the seeded bound stays loose for the
                                                        // whole scan (~4 survivors per tile) — profiles/r2_scorer_small_shard.md
constexpr int SM_KSPLIT = 1;                            // accumulation chains per tile (1: no split; the MMA
                                                        // probe shows dependent chains are not slower)
constexpr int SM_ACC_COLS = SM_KSPLIT * SM_ND;          // columns per buffer: partial sums x 32 docs
constexpr int SM_Q_COLS = SM_DIM;                       // A operand: one column per k element
constexpr int SM_TMEM_COLS = 512;
constexpr int SM_THREADS = 384;                         // warps: 0 TMA, 1 MMA (peer: forwarder), 2 TMEM alloc, 3 idle,
                                                        // 4-7 keepers (one thread per query), 8-11 screeners (ditto)
constexpr int SM_CAP = 128;                             // candidate slots per query
constexpr int SM_KEEP = 64;                             // a compaction leaves k..SM_KEEP entries
constexpr int SM_LIST_BYTES = SM_CAP * SM_MQ * 4;       // score lists ([slot][thread]): 64 KB
constexpr int SM_LISTI_BYTES = SM_CAP * SM_MQ * 2;      // id lists, uint16 CTA-local document numbers: 32 KB
constexpr int SM_MAX_TILES_PER_CTA = 65536 / SM_ND;     // so that a local document number fits 16 bits
constexpr int SM_SAMPLE_TOP = 4;                        // candidates a CTA publishes per query in the sample pass

__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
  if (v >= 0.f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

// maximum of 32 scores with the 3-input FMNMX3 of sm_100: 17 instructions, depth 4 (a pairwise tree: 31, depth 5) — on
// the single-warp dependent chain of the screener / keeper per 32-document accumulator half
__device__ __forceinline__ float max3f(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
template <typename T>
__device__ __forceinline__ float max_of_32(const T (&v)[32]) {
  auto f = [&](int i) -> float {
    if constexpr (sizeof(T) == 4 && !std::is_same<T, float>::value) return __uint_as_float((uint32_t)v[i]);
    else return (float)v[i];
  };
  float m[11];
#pragma unroll
  for (int j = 0; j < 10; ++j) m[j] = max3f(f(3 * j), f(3 * j + 1), f(3 * j + 2));
  m[10] = fmaxf(f(30), f(31));
  const float a = max3f(m[0], m[1], m[2]), b = max3f(m[3], m[4], m[5]), c = max3f(m[6], m[7], m[8]), d = fmaxf(m[9], m[10]);
  return fmaxf(max3f(a, b, c), d);
}

// order-preserving float <-> uint32 key
__device__ __forceinline__ uint32_t f2key(float f) {
  const uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key2f(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// Counts of list entries >= each of three pivots in ONE pass (slots >= cnt hold -inf).
// Two independent accumulator sets and 16 loads in flight: the pass is bound by shared-memory
// throughput, not by a dependent add chain.
__device__ __forceinline__ void count_ge3(const float* ls, float p1, float p2, float p3, int& c1, int& c2, int& c3) {
  int a1 = 0, a2 = 0, a3 = 0, b1 = 0, b2 = 0, b3 = 0;
#pragma unroll 8
  for (int e = 0; e < SM_CAP; e += 2) {
    const float v0 = ls[e * SM_MQ], v1 = ls[(e + 1) * SM_MQ];
    a1 += v0 >= p1; a2 += v0 >= p2; a3 += v0 >= p3;
    b1 += v1 >= p1; b2 += v1 >= p2; b3 += v1 >= p3;
  }
  c1 = a1 + b1; c2 = a2 + b2; c3 = a3 + b3;
}

// Per-thread compaction (called by a whole warp; lanes with cnt <= SM_KEEP only tag along).
// Keeps every entry >= pivot where k <= #kept <= SM_KEEP, or — when a plateau of equal scores
// straddles that window — the entries above the plateau plus its first (lowest-id) members
// up to exactly k.  Scores and ids both live in shared memory (an earlier version kept the ids
// in an L2 scratch: moving them cost ~45 k cycles per round under a saturated memory system).
// The pivot is searched with three value-space probes per pass (the
// interval shrinks 4x per pass; ~3 passes on score-like data) and, if that stalls, by
// bisection over the order-preserving integer keys, which always terminates.
// Returns the new count; tau = score bound for future candidates, strict = candidates must
// beat tau strictly (plateau case: later equal scores have larger ids and lose the tie).
__device__ __forceinline__ int thread_compact(float* ls, uint16_t* li, int cnt, int k, float& tau, bool& strict) {
  const bool need = cnt > SM_KEEP;
  for (int e = cnt; e < SM_CAP; ++e) ls[e * SM_MQ] = -INFINITY;      // pads never count
  float mn = INFINITY, mx = -INFINITY;
#pragma unroll 8
  for (int e = 0; e < SM_CAP; ++e) {
    const float v = ls[e * SM_MQ];
    if (e < cnt) { mn = fminf(mn, v); mx = fmaxf(mx, v); }
  }
  // invariant in key space: count(>= lo) > SM_KEEP, count(>= hi) < k
  uint32_t lo = f2key(mn), hi = f2key(mx) + 1u;
  float pivot = mn;
  bool plateau = false;
  bool done = !need;
  int iter = 0;
  while (__any_sync(0xffffffffu, !done)) {
    const float flo = key2f(lo), fhi = key2f(hi - 1u);
    float p1, p2, p3;
    uint32_t k1, k2, k3;
    const uint32_t kmid = lo + ((hi - lo) >> 1);
    p1 = flo + 0.25f * (fhi - flo); p2 = flo + 0.5f * (fhi - flo); p3 = flo + 0.75f * (fhi - flo);
    k1 = f2key(p1); k2 = f2key(p2); k3 = f2key(p3);
    const bool value_ok = iter < 10 && k1 > lo && k1 < k2 && k2 < k3 && k3 < hi;
    if (!value_ok) { k1 = k2 = k3 = kmid; p1 = p2 = p3 = key2f(kmid); }
    ++iter;
    int c1, c2, c3;
    count_ge3(ls, p1, p2, p3, c1, c2, c3);
    if (!done) {
      if (hi - lo <= 1u) { plateau = true; pivot = key2f(lo); done = true; }
      else if (c1 >= k && c1 <= SM_KEEP) { pivot = p1; done = true; }
      else if (c2 >= k && c2 <= SM_KEEP) { pivot = p2; done = true; }
      else if (c3 >= k && c3 <= SM_KEEP) { pivot = p3; done = true; }
      else {
        // counts are non-increasing in the pivot: tighten both ends
        if (c3 > SM_KEEP) lo = k3; else if (c2 > SM_KEEP) lo = k2; else if (c1 > SM_KEEP) lo = k1;
        if (c1 < k) hi = k1; else if (c2 < k) hi = k2; else if (c3 < k) hi = k3;
      }
    }
  }
  int w = cnt;
  if (need) {
    // stable in-place compaction; plateau: keep > pivot, then the first ties until k entries
    int n_gt = 0;
    if (plateau) { int d2, d3; const float pg = key2f(lo + 1u); count_ge3(ls, pg, pg, pg, n_gt, d2, d3); }
    int ties_left = plateau ? (k - n_gt) : 0;
    w = 0;
#pragma unroll 1
    for (int e0 = 0; e0 < SM_CAP; e0 += 8) {
      if (e0 >= cnt) break;
      float v[8];
      uint16_t id[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        v[j] = ls[(e0 + j) * SM_MQ];
        id[j] = li[(e0 + j) * SM_MQ];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        bool keep = false;
        if (e0 + j < cnt) {
          if (!plateau) keep = v[j] >= pivot;
          else if (v[j] > pivot) keep = true;
          else if (v[j] == pivot && ties_left > 0) { keep = true; --ties_left; }
        }
        if (keep) { ls[w * SM_MQ] = v[j]; li[w * SM_MQ] = id[j]; ++w; }
      }
    }
    tau = pivot;
    strict = plateau;
  }
  return w;
}

// Optional timeline trace of CTA (0,0): trace[role][tile] = clock64 at a pipeline event
// (roles: 0 producer issued, 1 MMA saw full, 2 MMA issued+committed, 3 epilogue saw acc_full,
// 4 epilogue released the accumulator).  Test/diagnostic only (ttr_debug_set_trace).
long long* g_score_trace = nullptr;
constexpr int SM_TRACE_TILES = 256;   // trace buffer: 8 roles x 256 tiles
#define SM_TRACE(role, it)                                                                      \
  do {                                                                                          \
    const int _ti = (it) - ((dbg & 128) ? 1000 : 0);                                            \
    if (trace && blockIdx.x == 0 && blockIdx.y == 0 && _ti >= 0 && _ti < SM_TRACE_TILES && lane == 0) \
      trace[(role) * SM_TRACE_TILES + _ti] = clock64();                                         \
  } while (0)

// Survivor histogram (main pass, seeded searches): SM_HBINS linear bins per query above the seeded bound.  Every
// candidate a keeper appends is counted (one fire-and-forget atomic) in the bin of its score; the lower edge of the
// highest bin whose suffix count reaches k is a lower bound on the final k-th best that tightens while ALL CTAs scan,
// not only when one CTA's own list fills.  On a small shard (233 tiles per CTA at 1.1 M documents) a list never fills:
// the bound stayed where the sample put it (top 0.1 %) and ~4 documents per tile survived to the keepers, which
// then set the tile rate (r2 trace: 1,320 cycles per tile against 1,000 of HBM time; CTA pairs 2,430 against 1,415).
constexpr int SM_HBINS = 64;
// bins of query q: [lo + b * w, lo + (b + 1) * w), the last one open-ended; w = 0 switches the histogram off
__device__ __forceinline__ float2 hist_params(float kth, float smax) {
  const bool ok = kth > -INFINITY && smax > kth && smax < INFINITY;
  // the final k-th best usually lies below the sample's maximum, sometimes above it: cover twice that span
  return ok ? make_float2(kth, 2.0f * (smax - kth) / (float)SM_HBINS) : make_float2(-INFINITY, 0.f);
}

__global__ void init_tau_kernel(float* tau, int32_t* qcount, int n, unsigned int* counters, unsigned int* hist,
                                float2* hpar) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  for (int j = i; j < n; j += gridDim.x * blockDim.x) { tau[j] = -INFINITY; qcount[j] = 0; hpar[j] = make_float2(-INFINITY, 0.f); }
  if (i < 2 && counters) counters[i] = 0u;
  for (int j = i; j < n * SM_HBINS; j += gridDim.x * blockDim.x) hist[j] = 0u;
}

// tau[q] = k-th best score of the sample pass (a valid lower bound on the final k-th best)
__global__ void seed_tau_kernel(float* tau, int32_t* qcount, const float* __restrict__ sample_s,
                                const int64_t* __restrict__ sample_i, int B, int n_pad, int k, float2* hpar) {
  int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q < n_pad) qcount[q] = 0;
  if (q < B && sample_i[(int64_t)q * k + (k - 1)] >= 0) {
    tau[q] = sample_s[(int64_t)q * k + (k - 1)];
    hpar[q] = hist_params(sample_s[(int64_t)q * k + (k - 1)], sample_s[(int64_t)q * k]);
  }
}

// Grid-barrier wait of the fused launch (all CTAs are co-resident: cooperative launch).  Bounded like the mbarrier
// waits: a protocol bug traps (launch failure) after ~4 s instead of hanging the GPU.
__device__ __forceinline__ void grid_wait(const unsigned int* counter, unsigned int expected) {
  for (unsigned int spin = 0; *reinterpret_cast<const volatile unsigned int*>(counter) < expected; ++spin) {
    __nanosleep(64);
    if (spin > (1u << 26)) __trap();
  }
}

// k-th largest of n floats in global memory (fused sample phase), computed by the 128 epilogue threads of a
// CTA (named barrier 2): keys staged in `smem_f` (n <= SM_CAP * SM_MQ), MSD radix select over the
// order-preserving 32-bit keys, 8 bits per round.  Returns -inf when fewer than k finite values exist.
// `scratch`: 258 words of shared memory nobody else uses meanwhile (the caller passes the idle id lists — static shared
// memory is down to its last few hundred bytes next to the 225 KB of ring + lists).
__device__ __forceinline__ float kth_largest_128(float* smem_f, uint32_t* scratch, const float* __restrict__ src, int n,
                                                 int k, int tid, float& vmax) {
  uint32_t* hist = scratch;
  uint32_t& s_prefix = scratch[256];
  uint32_t& s_rem = scratch[257];
  uint32_t* keys = reinterpret_cast<uint32_t*>(smem_f);
  int finite = 0;
  uint32_t kmax = 0u;
  // eight independent loads per thread in flight (148 slices x 4 values = 592 = 4.6 per thread: as a plain loop these
  // were five dependent L2 round trips in the middle of the grid-barrier phase, during which no CTA streams documents)
  for (int i0 = 0; i0 < n; i0 += 8 * SM_MQ) {
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = i0 + u * SM_MQ + tid;
      v[u] = i < n ? __ldcg(src + i) : -INFINITY;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = i0 + u * SM_MQ + tid;
      if (i < n) {
        const uint32_t key = f2key(v[u]);
        keys[i] = key;
        kmax = max(kmax, key);
        finite += v[u] > -INFINITY ? 1 : 0;
      }
    }
  }
  if (tid == 0) { s_prefix = 0u; s_rem = (uint32_t)k; }
  // count finite values across the 128 threads through the histogram array (hist[1]: the largest key)
  if (tid < 256 / 2) { hist[tid] = 0u; hist[tid + 128] = 0u; }
  ptx::named_bar_sync(2, SM_MQ);
  atomicAdd(&hist[0], (uint32_t)finite);
  atomicMax(&hist[1], kmax);
  ptx::named_bar_sync(2, SM_MQ);
  const bool enough = hist[0] >= (uint32_t)k;
  vmax = key2f(hist[1]);
  ptx::named_bar_sync(2, SM_MQ);
  if (!enough) return -INFINITY;
  for (int shift = 24; shift >= 0; shift -= 8) {
    hist[tid] = 0u; hist[tid + 128] = 0u;
    ptx::named_bar_sync(2, SM_MQ);
    const uint32_t prefix = s_prefix;
    const uint32_t hi_mask = shift == 24 ? 0u : (0xffffffffu << (shift + 8));
    for (int i = tid; i < n; i += SM_MQ) {
      const uint32_t key = keys[i];
      if ((key & hi_mask) == (prefix & hi_mask)) atomicAdd(&hist[(key >> shift) & 255u], 1u);
    }
    ptx::named_bar_sync(2, SM_MQ);
    if (tid < 32) {
      uint32_t h[8], own = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) { h[j] = hist[8 * tid + j]; own += h[j]; }
      uint32_t suf = own;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t v = __shfl_down_sync(0xffffffffu, suf, d);
        if (tid + d < 32) suf += v;
      }
      const uint32_t above = suf - own, rem = s_rem;
      __syncwarp();
      if (above < rem && rem <= suf) {
        uint32_t cum = above;
#pragma unroll
        for (int j = 7; j >= 0; --j) {
          if (cum < rem && rem <= cum + h[j]) {
            s_prefix = (prefix & hi_mask) | ((uint32_t)(8 * tid + j) << shift);
            s_rem = rem - cum;
          }
          cum += h[j];
        }
      }
    }
    ptx::named_bar_sync(2, SM_MQ);
  }
  return key2f(s_prefix);
}

// Scratch layout per CTA c = slice * n_qt + qt (thread ql = query inside the tile):
// Scratch layout (q = global query index, padded to 128 per tile; cap = n_slices * SM_KEEP):
//   out_s  [q][cap] fp32   candidate scores, appended by the CTAs in arrival order
//   out_i  [q][cap] i32    candidate ids
//   out_n  [q] i32         number of candidates (atomic cursor)
// A CTA only publishes candidates >= the global bound tau_g it sees when it finishes (k documents
// at or above that bound exist somewhere, so anything below cannot be in the top-k): a few
// hundred entries per query reach the merge kernel instead of n_slices * 64.
//
// SAMPLE = true is the threshold-seeding pass over the first few tiles of every slice: the epilogue
// thread keeps only its SM_SAMPLE_TOP best scores in registers (branch-free insertion, no lists,
// no compaction) and publishes those.  The k-th best of the union over all CTAs is a valid lower
// bound on the final k-th best (every published candidate is a real document), and it equals the
// exact k-th best of the sample unless one CTA holds more than SM_SAMPLE_TOP of the sample's top k
// (256-512 documents of ~38-76 k per CTA: mean 0.34 of the top 50; with 4 kept P(more) ~ 2e-4 per CTA, and
// then the bound is merely a little looser).  Keeping 4 instead of 8 halves the bubble and the number of
// documents that survive in SOME lane of the warp-uniform loop: 0.263 -> 0.246 ms per step on a 1.1 M shard.
//
// MODE 2 (one query tile, all CTAs co-resident — cooperative launch) fuses the two passes into one launch:
// every CTA first runs its sample tiles through the register top-4 path and writes them to a fixed slot of
// `samp` ([query][slice][8]); a grid-wide barrier; CTA q (q < B) radix-selects the k-th best of query q's
// n_slices * 8 sample scores into tau_g[q]; a second barrier; then the main pass over ALL tiles with the seeded
// bound.  The TMA and MMA warps simply keep running ahead into the main tiles while the barriers pass.  This
// removes a kernel launch with its ~30 us of fixed cost (TMEM allocation, query staging, pipeline fill, drain),
// the sample merge and the seed kernel: what limits an eighth-of-the-corpus shard (173 us of HBM time).
struct FusedArgs {
  float* samp;              // [128][n_slices * SM_SAMPLE_TOP] sample scores
  unsigned int* counters;   // two grid-barrier counters, zeroed before the launch
  int sample_tiles;         // tiles per CTA in the sample phase
  unsigned int* hist;       // [queries][SM_HBINS] survivor histogram of the main pass (all modes but the sample pass), or NULL
  float2* hpar;             // [queries] (lower edge of bin 0, bin width) — written by whoever seeds the bounds
};

// PAIR = true (query batches > 128): the two CTAs of a cluster own two adjacent 128-query tiles and the SAME document
// slice, and run ONE tcgen05.mma.cta_group::2 (M = 256) per k-step over 64-document tiles: each CTA stages its own
// queries in its own tensor memory and loads only ITS HALF (32 documents) of every tile; the tensor cores of both SMs
// read both halves.  Per SM that halves the document bytes delivered per unit of MMA work: with one CTA per query tile
// the L2 -> SM fabric saturated at ~7-8 TB/s (two CTAs fetching every tile: B = 256 ran at 0.55 of HBM, r2 traces);
// a pair fetches every tile once, so B = 256 is HBM-bound like B = 128.  Rank 0 (the leader) issues the MMAs; the
// TMA loads of both CTAs count their bytes on the leader's `full` barrier; tcgen05.commit multicasts `stage free` and
// `accumulator full` to both CTAs; the peer's epilogue warps release accumulators with remote arrives on the
// leader's `accumulator empty` barrier.
template <int MODE, bool PAIR>
__device__ __forceinline__ void
scorer_body(const float* __restrict__ Q, const CUtensorMap& map_d, int B, int64_t N,
            int k, int n_slices, float* __restrict__ tau_g, float* __restrict__ out_s,
            int32_t* __restrict__ out_i, int32_t* __restrict__ out_n, long long* __restrict__ trace,
            int dbg, FusedArgs fa, int seg_tiles, int cap) {
  constexpr bool SAMPLE = MODE == 1;
  constexpr bool FUSED = MODE == 2;
  constexpr int ND_T = PAIR ? 2 * SM_ND : SM_ND;       // documents per (pair-)tile == accumulator columns per buffer
  constexpr int HALVES = PAIR ? 2 : 1;                 // 32-column accumulator halves the epilogue reads per tile
  constexpr int SM_NACC = PAIR ? 4 : SM_NACC_MAX;      // accumulator buffers: 256 TMEM columns either way
  const uint32_t cta_rank = PAIR ? ptx::cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0u;
  // which 32 documents of a 64-document pair tile this CTA loads: B rows [0, 32) come from rank 0's shared memory,
  // rows [32, 64) from rank 1's (debug bit 26 swaps the assignment: bring-up switch for that convention)
  const int my_half = PAIR ? (int)(cta_rank ^ ((dbg >> 26) & 1u)) : 0;
  // How the leader learns that the PEER's half of a stage has landed.  Default: every CTA loads with the plain TMA form
  // onto its OWN `full` barrier and the peer's otherwise idle warp 1 forwards each completion with a remote arrive on the
  // leader's `peer_full` barrier.  Debug bit 28: the cta_group::2 TMA form counting both CTAs' bytes on the leader's
  // barrier — measured 2.6x slower (r2 traces: 7-9 k cycles per 32 KB stage at half the HBM rate, against 2.9 k with plain
  // loads), so it is kept only as the A/B switch.
  const bool tma_2cta = PAIR && (dbg & (1 << 28));
  extern __shared__ unsigned char smem_raw[];
  unsigned char* base = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* ring = base;                                      // [stages][SM_KB][32 rows][128 B]
  float* list_s = reinterpret_cast<float*>(ring + SM_STAGES * SM_STAGE_BYTES);   // [SM_CAP][128]
  uint16_t* list_i = reinterpret_cast<uint16_t*>(reinterpret_cast<unsigned char*>(list_s) + SM_LIST_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(list_i) + SM_LISTI_BYTES);
  uint64_t* empty_bar = full_bar + SM_STAGES;
  uint64_t* acc_full = empty_bar + SM_STAGES;
  uint64_t* acc_empty = acc_full + SM_NACC_MAX;
  uint64_t* peer_full = acc_empty + SM_NACC_MAX;       // pair, leader only: "the peer's half of stage s has landed"
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(peer_full + SM_STAGES);
  __shared__ uint64_t probe_bar;             // dbg & 64: the MMA warp waits for its own commit (timing experiment)
  // Screening (default; debug bit 29 switches it off).  The epilogue used to be ONE warp per scheduler doing everything
  // for its 32 queries: wait, tcgen05.ld, filter, list upkeep — a ~1,500-cycle dependent chain per 64-document tile
  // that nothing overlapped, the bound of every batch above 128 queries once the CTA pairs had fixed the operand feed.
  // Now a second warp per lane quarter (a "screener", warps 8-11) reads every accumulator first, reduces the thread's
  // scores to their maximum and votes against the queries' current thresholds; tiles in which no query of the quarter
  // has a survivor (~3 of 4 once the bounds are seeded) are released right there.  Only the others are queued for the
  // quarter's "keeper" warp (4-7), which owns the candidate lists and runs the unchanged filter / compaction code on
  // them, in tile order.  Thresholds flow keeper -> screener through `thr_sh` (stale = looser = safe).
  __shared__ float thr_sh[SM_MQ];            // per query: a score below this cannot enter the list
  __shared__ int hit_q[4][16];               // per lane quarter: tiles the keeper must process (<= SM_NACC pending)
  __shared__ int hit_wr[4];                  // entries written to hit_q
  __shared__ int scr_tiles[4];               // tiles the screener has classified
  __shared__ int go_main[4];                 // fused launch: the keeper has published the seeded bounds
  __shared__ float tgl_sh[SM_MQ];            // per query: the bound warps 2-3 read off the global survivor histogram / tau_g
  __shared__ int keepers_done;               // keeper warps that have published their lists (the sweepers' exit signal)
  const bool screen_on = !(dbg & (1 << 29));
  const bool hist_on = screen_on && MODE != 1 && fa.hist != nullptr && dbg >= 0;      // debug bit 31: histogram bound off

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#define SM_MARK(slot)                                                                                   \
  do {                                                                                                  \
    if (trace && blockIdx.x == 0 && blockIdx.y == 0) trace[(slot) * SM_TRACE_TILES + SM_TRACE_TILES - 1] = clock64(); \
  } while (0)
  if (threadIdx.x == 0) SM_MARK(0);          // kernel entry (slots use the last column of the trace rows)
  // dbg bit 20: instead of the tile timeline, rows 5/6 of the trace receive every CTA's entry / exit globaltimer (ns)
  if (trace && (dbg & (1 << 20)) && threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y < SM_TRACE_TILES) {
    long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    trace[5 * SM_TRACE_TILES + blockIdx.y] = gt;
  }
  const int qt = blockIdx.x;                 // query tile
  const int slice = blockIdx.y;              // document slice
  const int q0 = qt * SM_MQ;
  const int64_t n_tiles = ceil_div64(N, (int64_t)ND_T);
  // this CTA's tile sequence: [FUSED: its first S tiles once more in front,] then tiles slice, slice + n_slices, ...
  const int n_mine = slice < n_tiles ? (int)((n_tiles - slice + n_slices - 1) / n_slices) : 0;
  const int S = FUSED ? fa.sample_tiles : 0;
  const int n_seq = n_mine + S;
  auto tile_at = [&](int it) -> int64_t { return (int64_t)slice + (int64_t)(it < S ? it : it - S) * n_slices; };

  if (threadIdx.x == 0) {
    for (int s = 0; s < SM_STAGES; ++s) {
      ptx::mbar_init(full_bar + s, 1); ptx::mbar_init(empty_bar + s, 1); ptx::mbar_init(peer_full + s, 1);
    }
    // accumulator release: the four epilogue warps of this CTA (+ the four of the peer, on the leader's barrier)
    for (int b = 0; b < SM_NACC; ++b) { ptx::mbar_init(acc_full + b, 1); ptx::mbar_init(acc_empty + b, PAIR ? 8 : 4); }
    ptx::mbar_init(&probe_bar, 1);
    ptx::fence_mbar_init();
    for (int w = 0; w < 4; ++w) { hit_wr[w] = 0; scr_tiles[w] = 0; go_main[w] = 0; }
    keepers_done = 0;
  }
  if (threadIdx.x < SM_MQ) tgl_sh[threadIdx.x] = -INFINITY;
  if (warp == 2) {
    if (PAIR) { ptx::tmem_alloc_2cta(tmem_slot, SM_TMEM_COLS); ptx::tmem_relinquish_2cta(); }
    else { ptx::tmem_alloc(tmem_slot, SM_TMEM_COLS); ptx::tmem_relinquish(); }
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_acc = tmem_base + SM_Q_COLS;

  // ---- stage the query tile into tensor memory: thread (warp 4+w, lane) owns TMEM lane 32w+lane
  // (r2: splitting the columns over the keeper and the screener warp of a quarter and letting the TMA producer start
  // before this barrier did not help — 18.1 k instead of 13.8 k cycles until "queries staged", kernel 322.9 k vs 320.0 k)
  if (warp >= 4 && warp < 8) {
    const int qw = warp - 4;
    const int q = q0 + qw * 32 + lane;
    // initial screening threshold: the seeded global bound (main pass), "anything" while a sample top-4 is empty,
    // +inf for the padding queries of the last tile
    thr_sh[qw * 32 + lane] = q >= B ? INFINITY : (MODE == 0 ? __ldcg(tau_g + q) : -3.402823466e38f);
    // Two rounds of 128 columns.  Loads are COALESCED — instruction i of a warp reads the 512-byte segment of row i of
    // its quarter, lane = 16-byte chunk — and transposed through a warp-private scratch in the (still idle) ring so that
    // lane r ends up with row r: a thread reading its own 1 KB row issued 32-line requests per instruction, 8,192 line
    // requests per CTA, and the prologue was bound by their throughput (13.8 k cycles to "queries staged").
    unsigned char* scratch = ring + qw * (32 * 528);                 // 32 rows x (512 B + 16 B pad): conflict-free both ways
#pragma unroll 1
    for (int c0 = 0; c0 < SM_DIM; c0 += 128) {
      float4 x[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int qrow = q0 + qw * 32 + i;
        x[i] = (qrow < B) ? __ldg(reinterpret_cast<const float4*>(Q + (int64_t)qrow * SM_DIM + c0) + lane)
                          : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int i = 0; i < 32; ++i) *reinterpret_cast<float4*>(scratch + i * 528 + lane * 16) = x[i];
      __syncwarp();
#pragma unroll
      for (int v = 0; v < 32; ++v) x[v] = *reinterpret_cast<const float4*>(scratch + lane * 528 + v * 16);
      __syncwarp();
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint32_t r[32];
#pragma unroll
        for (int v = 0; v < 8; ++v) {
          r[4 * v + 0] = __float_as_uint(round_tf32(x[8 * g + v].x));
          r[4 * v + 1] = __float_as_uint(round_tf32(x[8 * g + v].y));
          r[4 * v + 2] = __float_as_uint(round_tf32(x[8 * g + v].z));
          r[4 * v + 3] = __float_as_uint(round_tf32(x[8 * g + v].w));
        }
        ptx::tmem_st_32x32(tmem_base + ((uint32_t)(qw * 32) << 16) + c0 + 32 * g, r);
      }
    }
    ptx::tmem_st_wait();
    ptx::fence_proxy_async_smem();           // the scratch is ring space: generic writes before the TMA (async proxy) fills it
  }
  ptx::tc_fence_before_sync();
  // pair: the leader's MMAs read the PEER's staged queries and signal its barriers too -> cluster-wide barrier
  if (PAIR) ptx::cluster_sync(); else __syncthreads();
  ptx::tc_fence_after_sync();

  if (threadIdx.x == 0) SM_MARK(1);          // prologue done (TMEM allocated, queries staged)
  if (warp == 0) {
    // ===== TMA producer (whole warp runs the loop; one elected lane issues) =====
    if (ptx::elect_one()) ptx::prefetch_tensormap(&map_d);
    for (int it = 0; it < n_seq; ++it) {
      const int64_t t = tile_at(it);
      const int s = it % SM_STAGES;
      const uint32_t ph = (uint32_t)(it / SM_STAGES) & 1u;
      ptx::mbar_wait(empty_bar + s, ph ^ 1u);
      unsigned char* dst = ring + s * SM_STAGE_BYTES;
      const int32_t d0 = (int32_t)(t * ND_T) + my_half * SM_ND;
      if (ptx::elect_one()) {
        if (tma_2cta) {
          // both CTAs' bytes are counted on the LEADER's barrier (one arrival: the leader's expect_tx for 2 x 32 KB; a
          // peer load that lands first only drives the transaction count negative until then)
          if (leader) ptx::mbar_arrive_expect_tx(full_bar + s, 2 * SM_STAGE_BYTES);
          const uint32_t bar0 = ptx::mapa(ptx::smem_u32(full_bar + s), 0u);
#pragma unroll
          for (int kb = 0; kb < SM_KB; ++kb)
            ptx::tma_load_2d_2cta(dst + kb * SM_D_KB_BYTES, &map_d, kb * 32, d0, bar0);
        } else {
          ptx::mbar_arrive_expect_tx(full_bar + s, SM_STAGE_BYTES);
#pragma unroll
          for (int kb = 0; kb < SM_KB; ++kb)
            ptx::tma_load_2d(dst + kb * SM_D_KB_BYTES, &map_d, kb * 32, d0, full_bar + s);
        }
      }
      __syncwarp();
      SM_TRACE(0, it);
    }
  } else if (warp == 1 && leader) {
    // ===== MMA issuer (whole warp runs the loop; one elected lane issues; pair: the leader CTA only) =====
    constexpr uint32_t idesc = ptx::make_idesc_tf32(PAIR ? 2 * SM_MQ : SM_MQ, ND_T);
    for (int it = 0; it < n_seq; ++it) {
      const int64_t t = tile_at(it);
      const int s = it % SM_STAGES;
      const uint32_t ph = (uint32_t)(it / SM_STAGES) & 1u;
      const int buf = it % SM_NACC;
      const uint32_t aph = (uint32_t)(it / SM_NACC) & 1u;
      ptx::mbar_wait(acc_empty + buf, aph ^ 1u);
      ptx::mbar_wait(full_bar + s, ph);
      if (PAIR && !tma_2cta) ptx::mbar_wait(peer_full + s, ph);
      ptx::tc_fence_after_sync();
      SM_TRACE(1, it);
      const uint32_t d_addr = ptx::smem_u32(ring + ((dbg & 16) ? 0 : s) * SM_STAGE_BYTES);
      const uint32_t d_tmem = tmem_acc + buf * ND_T;
      const uint64_t b_desc0 = ptx::make_kmajor_sw128_desc(d_addr);
      if (ptx::elect_one()) {
        // Measured (profiles/r1_score_topk_mma_v3_trace_b16.txt vs v4): one accumulation chain
        // per tile runs at ~2950 cycles per tile, four interleaved chains at <= 1850, so the
        // K = 256 reduction is split into SM_KSPLIT partial accumulators (k-steps
        // [c*8, c*8+8) -> accumulator c) issued round-robin; the epilogue adds them.
#pragma unroll
        for (int kk = 0; kk < 32 / SM_KSPLIT; ++kk) {
#pragma unroll
          for (int c = 0; c < SM_KSPLIT; ++c) {
            const int ks = c * (32 / SM_KSPLIT) + kk;              // k-step 0..31 (8 floats each)
            // k-block (ks>>2) starts (ks>>2)*4 KB further (>>4 in descriptor units); 32 B per k-step inside it
            const uint64_t b_desc = b_desc0 + (uint64_t)(((dbg & 32) ? 0 : (ks >> 2)) * (SM_D_KB_BYTES >> 4) + 2 * (ks & 3));
            if (PAIR) ptx::mma_tf32_ts_2cta(d_tmem, tmem_base + ks * 8, b_desc, idesc, kk != 0);
            else ptx::mma_tf32_ts(d_tmem + c * SM_ND, tmem_base + ks * 8, b_desc, idesc, kk != 0);
          }
        }
        if (PAIR) {
          ptx::mma_commit_2cta(empty_bar + s, 3);       // both CTAs' producers may refill this stage
          ptx::mma_commit_2cta(acc_full + buf, 3);      // both CTAs' epilogues may read this accumulator
        } else {
          ptx::mma_commit(empty_bar + s);
          ptx::mma_commit(acc_full + buf);
        }
        if (dbg & 64) ptx::mma_commit(&probe_bar);
      }
      __syncwarp();
      SM_TRACE(2, it);
      if (dbg & 64) {
        ptx::mbar_wait(&probe_bar, (uint32_t)it & 1u);
        SM_TRACE(5, it);
      }
    }
  } else if (PAIR && warp == 1 && !tma_2cta) {
    // ===== peer CTA: forward "my half of stage s has landed" to the leader's MMA issuer =====
    for (int it = 0; it < n_seq; ++it) {
      const int s = it % SM_STAGES;
      const uint32_t ph = (uint32_t)(it / SM_STAGES) & 1u;
      ptx::mbar_wait(full_bar + s, ph);
      if (ptx::elect_one()) ptx::mbar_arrive_cluster(ptx::mapa(ptx::smem_u32(peer_full + s), 0u));
      __syncwarp();
    }
  } else if ((warp == 2 || warp == 3) && hist_on) {
    // ===== bound sweepers: the two otherwise idle warps turn the global survivor histogram into per-query bounds =====
    // Warp w sweeps queries [64 (w - 2), +64), four at a time: lane l reads bins 2l, 2l + 1 of a query (one coalesced
    // 256-byte row), a suffix sum over the lanes gives the number of counted documents at or above every bin edge, and
    // the highest edge with >= k of them — or tau_g, whichever is larger — goes to tgl_sh for the keepers and screeners.
    volatile int* go_v = go_main;
    volatile int* done_v = &keepers_done;
    volatile float* tgl_v = tgl_sh;
    if (FUSED) {
      for (unsigned spin = 0; go_v[0] == 0 && *done_v < 4; ++spin) {
        __nanosleep(256);
        if (spin > (1u << 24)) __trap();
      }
    }
    const int qb0 = 64 * (warp - 2);
    bool stop = false;
    // The bound rises like the logarithm of the scanned fraction, so sweeps are spaced geometrically: the next one
    // starts when the scan is 1/8 older (at least 2 us later).  Back-to-back sweeps of all CTAs were ~1 TB/s of L2
    // reads for nothing on a full-corpus scan.
    const long long t_start = ptx::globaltimer_ns();
    while (!stop) {
      const long long t_now = ptx::globaltimer_ns();
      const long long t_next = t_now + max(2000ll, (t_now - t_start) >> 3);
#pragma unroll 1
      for (int g = 0; g < 64 && !stop; g += 4) {
        uint2 c[4];
        float2 hp[4];
        float tq[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int q = q0 + qb0 + g + u;                    // < n_qt * 128: the scratch is padded to whole query tiles
          c[u] = __ldcg(reinterpret_cast<const uint2*>(fa.hist + (size_t)q * SM_HBINS) + lane);
          hp[u] = __ldcg(fa.hpar + q);
          tq[u] = __ldcg(tau_g + q);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          uint32_t suf = c[u].x + c[u].y;                    // documents counted in this lane's two bins ...
#pragma unroll
          for (int d = 1; d < 32; d <<= 1) {
            const uint32_t v = __shfl_down_sync(0xffffffffu, suf, d);
            if (lane + d < 32) suf += v;                     // ... and in all higher bins
          }
          const int cand = (suf - c[u].x >= (uint32_t)k) ? 2 * lane + 1 : (suf >= (uint32_t)k ? 2 * lane : -1);
          const int best = __reduce_max_sync(0xffffffffu, cand);
          const float bound = (best >= 0 && hp[u].y > 0.f) ? fmaf((float)best, hp[u].y, hp[u].x) : -INFINITY;
          if (lane == 0) tgl_v[qb0 + g + u] = fmaxf(bound, tq[u]);
        }
        stop = *done_v >= 4;
      }
      while (!stop && ptx::globaltimer_ns() < t_next) {
        __nanosleep(500);
        stop = *done_v >= 4;
      }
    }
  } else if (warp >= 8) {
    // ===== screener: one thread per query; classifies every tile, releases the ones without a survivor =====
    const int qw = warp - 8;
    const int ql = qw * 32 + lane;
    const int q = q0 + ql;
    const bool q_valid = q < B;
    volatile float* thr_v = thr_sh;
    volatile int* hit_wr_v = hit_wr;
    volatile int* scr_v = scr_tiles;
    volatile int* go_v = go_main;
    float tgp = (q_valid && MODE == 0) ? __ldcg(tau_g + q) : (q_valid ? -INFINITY : INFINITY);
    float tgp_pending = tgp;
    int wr = 0;
    // Sample phase: the screener alone serves it.  Per tile it keeps the thread's MAXIMUM (the FMNMX tree it computes
    // anyway) and the best SM_SAMPLE_TOP tile maxima of the phase; those are published instead of an exact per-thread
    // top-4.  Every published value is the score of a distinct real document, so the k-th best of the union over all
    // CTAs is still a valid lower bound on the final k-th best (a tile holding two of the sample's top k contributes
    // one: the bound is marginally looser).  The exact top-4 insertion it replaces ran a 32-iteration warp-uniform
    // loop per tile while the thresholds were loose: 5-17 k cycles per tile, 40-70 us per search on small shards.
    float smp_s[SM_SAMPLE_TOP];
    int32_t smp_i[SM_SAMPLE_TOP];
#pragma unroll
    for (int e = 0; e < SM_SAMPLE_TOP; ++e) { smp_s[e] = -INFINITY; smp_i[e] = -1; }
    auto publish_sample_fused = [&]() {
      float* mine = fa.samp + ((size_t)(q0 + ql) * n_slices + slice) * SM_SAMPLE_TOP;
#pragma unroll
      for (int e = 0; e < SM_SAMPLE_TOP; ++e) mine[e] = (q_valid && smp_i[e] >= 0) ? smp_s[e] : -INFINITY;
      __threadfence();
    };
    for (int it = 0; screen_on && it < n_seq; ++it) {
      if (FUSED && it == S) {
        publish_sample_fused();
        __syncwarp();
        if (lane == 0) {
          __threadfence_block();
          scr_v[qw] = S;                       // the keeper may start the grid-barrier phase
        }
        // the keeper publishes the seeded bounds after the grid barriers; main tiles must not be judged by the
        // sample thresholds (those say "beats my 4th best", not "can be in the top k")
        for (unsigned spin = 0; go_v[qw] == 0; ++spin) {
          __nanosleep(64);
          if (spin > (1u << 26)) __trap();
        }
        tgp = q_valid ? __ldcg(tau_g + q) : INFINITY;
        tgp_pending = tgp;
      }
      const bool smp = SAMPLE || (FUSED && it < S);
      if ((it & 7) == 0 && q_valid && !smp) {
        tgp = fmaxf(tgp, tgp_pending);
        tgp_pending = __ldcg(tau_g + q);
      }
      const int64_t t = tile_at(it);
      const int buf = it % SM_NACC;
      ptx::mbar_wait(acc_full + buf, (uint32_t)(it / SM_NACC) & 1u);
      ptx::tc_fence_after_sync();
      float mx = -INFINITY;
      {
        const uint32_t tcol = tmem_acc + ((uint32_t)(qw * 32) << 16) + buf * ND_T;
        uint32_t r[HALVES][32];
#pragma unroll
        for (int h = 0; h < HALVES; ++h) ptx::tmem_ld_32x32(tcol + h * SM_ND, r[h]);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int h = 0; h < HALVES; ++h) {
          mx = fmaxf(mx, max_of_32(r[h]));
        }
      }
      const bool tail = (t + 1) * ND_T > N;    // zero-filled rows beyond N (score 0 is not a document's score)
      if (smp) {
        float cs = (q_valid && !tail) ? mx : -INFINITY;
        int32_t ci = (int32_t)(t * ND_T);      // a distinct id per tile (the merge of the sample pass wants distinct keys)
#pragma unroll
        for (int e = 0; e < SM_SAMPLE_TOP; ++e) {
          const bool up = cs > smp_s[e];
          const float ts = smp_s[e];
          const int32_t ti = smp_i[e];
          smp_s[e] = up ? cs : ts;
          smp_i[e] = up ? ci : ti;
          cs = up ? ts : cs;
          ci = up ? ti : ci;
        }
      }
      if (!smp) tgp = fmaxf(tgp, *(volatile float*)(tgl_sh + ql));
      const float thr = fmaxf(thr_v[ql], tgp);
      // the keeper masks the zero-filled rows of the last tile, so it always gets that tile
      const bool hit = !smp && (__any_sync(0xffffffffu, mx >= thr) || tail);
      if (!hit) {
        ptx::tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) {
          if (!PAIR || leader) ptx::mbar_arrive(acc_empty + buf);
          else ptx::mbar_arrive_cluster(ptx::mapa(ptx::smem_u32(acc_empty + buf), 0u));
        }
      } else {
        if (lane == 0) {
          hit_q[qw][wr & 15] = it;
          __threadfence_block();
          hit_wr_v[qw] = wr + 1;
        }
        ++wr;
      }
      if (lane == 0 && !(FUSED && it + 1 == S)) {      // "sample done" is signalled after the sample values are written
        __threadfence_block();
        scr_v[qw] = it + 1;
      }
    }
    if (screen_on && SAMPLE && q_valid) {            // separate sample pass: tile maxima -> candidate scratch
      int n_keep = 0;
#pragma unroll
      for (int e = 0; e < SM_SAMPLE_TOP; ++e) n_keep += smp_i[e] >= 0 ? 1 : 0;
      size_t ob = (size_t)q * (size_t)cap + (size_t)atomicAdd(out_n + q, n_keep);
#pragma unroll
      for (int e = 0; e < SM_SAMPLE_TOP; ++e)
        if (smp_i[e] >= 0) { out_s[ob] = smp_s[e]; out_i[ob] = smp_i[e]; ++ob; }
    }
  } else if (warp >= 4) {
    // ===== keeper: one thread per query; owns the candidate list =====
    const int qw = warp - 4;
    const int ql = qw * 32 + lane;             // query inside the tile == TMEM lane == list column
    const int q = q0 + ql;
    const bool q_valid = q < B;
    float* ls = list_s + ql;
    uint16_t* li = list_i + ql;   // CTA-local document number: it * 32 + j
    float top_s[SM_SAMPLE_TOP];   // SAMPLE: best scores so far, descending
    int32_t top_i[SM_SAMPLE_TOP];
#pragma unroll
    for (int e = 0; e < SM_SAMPLE_TOP; ++e) { top_s[e] = -INFINITY; top_i[e] = -1; }
    float tau = -INFINITY;                     // own bound on the k-th best
    bool strict = false;
    int cnt = 0;
    int it = 0;
    int seg_base = 0;                          // first main-pass tile of the current id segment (see `publish`)
    // Global lower bound on the k-th best.  An L2 read under a saturated memory system costs
    // microseconds, so it is refreshed every 8 tiles and consumed one refresh later.
    float tg = (q_valid && !FUSED) ? __ldcg(tau_g + q) : (q_valid ? -INFINITY : INFINITY);   // seeded by the sample pass (or -inf)
    float tg_pending = tg;
    // survivor histogram of this thread's query: lower edge of bin 0, bin width (0 = off), 1 / width
    float h_lo = 0.f, h_w = 0.f, h_iw = 0.f;
    unsigned int* h_row = hist_on ? fa.hist + (size_t)q * SM_HBINS : nullptr;
    auto load_hist_params = [&]() {
      if (hist_on && q_valid) {
        const float2 hp = __ldcg(fa.hpar + q);
        h_lo = hp.x; h_w = hp.y; h_iw = hp.y > 0.f ? 1.0f / hp.y : 0.f;
      }
    };
    if (!FUSED) load_hist_params();
    // one tile of the epilogue; `smp` (compile-time) selects the register top-4 path of the sample phase.  Two
    // instantiations instead of a runtime flag: with the flag in the loop the main pass ran 10 % slower.
    // Filter one 32-score accumulator half: documents d0 .. d0 + 31, list ids lid0 + j.
    auto filter32 = [&](const float (&sc32)[32], const int64_t d0, const int lid0, auto smp_tag) {
      constexpr bool smp = decltype(smp_tag)::value;
      // one threshold, one compare per score: "strictly above tau" == ">= next float above tau"
      const float thr = smp ? key2f(f2key(top_s[SM_SAMPLE_TOP - 1]) + 1u)         // must beat the weakest kept score
                            : fmaxf(tg, strict ? key2f(f2key(tau) + 1u) : tau);
      // Fast reject: the maximum of the thread's 32 scores (a 5-level FMNMX tree, ~25 issue slots) against the
      // threshold, one vote per warp.  After the sample pass seeded the bounds only ~1 tile in 4 holds a survivor
      // for ANY of a warp's 32 queries (r2 traces: the per-score mask build + OR-reduce + compaction vote below
      // cost ~610 of the epilogue's ~1100 cycles per tile whether or not anything survived).
      const bool tail = d0 + 32 > N;          // documents >= N are zero-filled by TMA and must not count
      bool skip = false;
      if (!(dbg & (1 << 25))) {
        skip = !tail && !__any_sync(0xffffffffu, max_of_32(sc32) >= thr);
      }
      if (skip) return;
      // Branch-free filter -> per-lane bit mask of surviving documents.  (A short-circuit
      // condition compiles to a branch per score: ~45 cycles of resolve latency each with one
      // warp per scheduler, 1300-1600 cycles per tile; profiles/r1_score_topk_mma_v4_trace_*.)
      uint32_t mask = 0;
#pragma unroll
      for (int j = 0; j < 32; ++j) mask |= (sc32[j] >= thr ? 1u : 0u) << j;
      if (tail) mask &= (N - d0 >= 32) ? 0xffffffffu : (N <= d0 ? 0u : ((1u << (int)(N - d0)) - 1u));
      // Visit only the documents that survive in SOME lane (warp-uniform loop over the OR of
      // the lane masks, usually 0-2 bits): traversing 32 conditional regions costs ~1500 cycles
      // per tile even when nothing is appended (profiles/r1_score_topk_mma_v4_trace_*).
      uint32_t any_mask = __reduce_or_sync(0xffffffffu, mask);
      while (any_mask) {
        const int j = __ffs(any_mask) - 1;
        any_mask &= any_mask - 1u;
        float scj = 0.f;
        switch (j) {   // j is warp-uniform: a uniform jump selects the register
#define SM_CASE(J) case J: scj = sc32[J]; break;
          SM_CASE(0) SM_CASE(1) SM_CASE(2) SM_CASE(3) SM_CASE(4) SM_CASE(5) SM_CASE(6) SM_CASE(7)
          SM_CASE(8) SM_CASE(9) SM_CASE(10) SM_CASE(11) SM_CASE(12) SM_CASE(13) SM_CASE(14) SM_CASE(15)
          SM_CASE(16) SM_CASE(17) SM_CASE(18) SM_CASE(19) SM_CASE(20) SM_CASE(21) SM_CASE(22) SM_CASE(23)
          SM_CASE(24) SM_CASE(25) SM_CASE(26) SM_CASE(27) SM_CASE(28) SM_CASE(29) SM_CASE(30) SM_CASE(31)
#undef SM_CASE
        }
        if (smp) {
          // bubble the candidate through the sorted registers; lanes whose score did not pass carry -inf
          float cs = ((mask >> j) & 1u) ? scj : -INFINITY;
          int32_t ci = (int32_t)(d0 + j);
#pragma unroll
          for (int e = 0; e < SM_SAMPLE_TOP; ++e) {
            const bool up = cs > top_s[e];
            const float ts = top_s[e];
            const int32_t ti = top_i[e];
            top_s[e] = up ? cs : ts;
            top_i[e] = up ? ci : ti;
            cs = up ? ts : cs;
            ci = up ? ti : ci;
          }
        } else if ((mask >> j) & 1u) {   // cnt <= SM_CAP - SM_ND before every half, so 32 free slots exist
          ls[cnt * SM_MQ] = scj;
          li[cnt * SM_MQ] = (uint16_t)(lid0 + j);
          ++cnt;
          if (h_w > 0.f) {
            // bin b holds scores >= fmaf(b, w, lo) — the very expression the sweepers evaluate, so rounding in the
            // index arithmetic can only put a document one bin too LOW (a looser, still valid bound)
            int b = min(max((int)((scj - h_lo) * h_iw), 0), SM_HBINS - 1);
            if (scj < fmaf((float)b, h_w, h_lo)) --b;
            if (b >= 0) atomicAdd(h_row + b, 1u);
          }
        }
      }
      if (!smp && __any_sync(0xffffffffu, cnt > SM_CAP - SM_ND)) {
        __syncwarp();
        cnt = thread_compact(ls, li, cnt, k, tau, strict);
        if (q_valid && cnt >= k) atomic_max_float(tau_g + q, tau);
      }
    };
    // one (pair-)tile of the epilogue; `smp` (compile-time) selects the register top-4 path of the sample phase.  Two
    // instantiations instead of a runtime flag: with the flag in the loop the main pass ran 10 % slower.
    auto tile_body = [&](const int it, auto smp_tag) {
      constexpr bool smp = decltype(smp_tag)::value;
      const int64_t t = tile_at(it);
      const int buf = it % SM_NACC;
      const uint32_t aph = (uint32_t)(it / SM_NACC) & 1u;
      if ((screen_on || (it & 7) == 0) && q_valid && !smp) {
        tg = fmaxf(fmaxf(tg, tg_pending), *(volatile float*)(tgl_sh + ql));
        tg_pending = __ldcg(tau_g + q);
      }
      ptx::mbar_wait(acc_full + buf, aph);
      ptx::tc_fence_after_sync();
      if (qw == 0) SM_TRACE(3, it);
      float sc[HALVES][32];
      {
        const uint32_t tcol = tmem_acc + ((uint32_t)(qw * 32) << 16) + buf * ND_T;
#pragma unroll
        for (int h = 0; h < HALVES; ++h) {
          uint32_t r[32];
          if (!(dbg & 8)) {
            ptx::tmem_ld_32x32(tcol + h * SM_ND, r);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) r[j] = 0;
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) sc[h][j] = __uint_as_float(r[j]);
        }
        if (!(dbg & 8)) ptx::tmem_ld_wait();
      }
      ptx::tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) {
        // the accumulator may be overwritten: tell the MMA issuer (pair: it lives in the leader CTA)
        if (!PAIR || leader) ptx::mbar_arrive(acc_empty + buf);
        else ptx::mbar_arrive_cluster(ptx::mapa(ptx::smem_u32(acc_empty + buf), 0u));
      }
      if (qw == 0) SM_TRACE(4, it);
#pragma unroll
      for (int h = 0; h < HALVES; ++h)
        filter32(sc[h], t * ND_T + h * SM_ND, (it - S - seg_base) * ND_T + h * SM_ND, smp_tag);
      if (qw == 0) SM_TRACE(6, it);
      if (screen_on) {
        // what the screener compares against from now on
        const float thr_now = smp ? key2f(f2key(top_s[SM_SAMPLE_TOP - 1]) + 1u)
                                  : fmaxf(tg, strict ? key2f(f2key(tau) + 1u) : tau);
        *(volatile float*)(thr_sh + ql) = q_valid ? thr_now : INFINITY;
      }
      if (qw == 0) SM_TRACE(7, it);
    };
    // Publish the list: at most SM_KEEP candidates at or above the global bound go to the per-query
    // global scratch (ids become global document numbers), then the list restarts empty.  Called at the
    // end, and every `seg_tiles` main-pass tiles in between: a list entry numbers its document with 16
    // bits relative to the segment start, so a CTA may scan any number of tiles (one wave of CTAs for
    // every query-tile count, instead of n_tiles / 2048 slices in several waves).  tau / strict stay
    // valid across the reset: the k documents that justified them have been published.
    auto publish = [&]() {
      __syncwarp();
      if (__any_sync(0xffffffffu, cnt > SM_KEEP)) cnt = thread_compact(ls, li, cnt, k, tau, strict);
      if (q_valid) {
        const float tgf = fmaxf(__ldcg(tau_g + q), *(volatile float*)(tgl_sh + ql));
        int n_keep = 0;
        for (int e = 0; e < SM_KEEP; ++e) n_keep += (e < cnt && ls[e * SM_MQ] >= tgf) ? 1 : 0;
        size_t ob = (size_t)q * (size_t)cap + (size_t)atomicAdd(out_n + q, n_keep);
        for (int e = 0; e < SM_KEEP; ++e)
          if (e < cnt && ls[e * SM_MQ] >= tgf) {
            const int lid = li[e * SM_MQ];                     // local -> global: tile = slice + main_tile * n_slices
            out_s[ob] = ls[e * SM_MQ];
            out_i[ob] = (int32_t)(((int64_t)slice + (int64_t)(seg_base + lid / ND_T) * n_slices) * ND_T + lid % ND_T);
            ++ob;
          }
      }
      cnt = 0;
    };
    auto main_tile = [&](const int it2) {
      while (it2 - S - seg_base >= seg_tiles) {   // warp-uniform
        publish();
        seg_base += seg_tiles;
      }
      tile_body(it2, std::false_type{});
    };
    // screening on: process the tiles the quarter's screener queued, until it has classified tiles [0, end)
    volatile int* hit_wr_v = hit_wr;
    volatile int* scr_v = scr_tiles;
    int rd = 0;
    auto drain = [&](const int end, auto smp_tag) {
      for (unsigned spin = 0;;) {
        if (rd < hit_wr_v[qw]) {
          const int it2 = *(volatile int*)&hit_q[qw][rd & 15];
          ++rd;
          if (decltype(smp_tag)::value) tile_body(it2, std::true_type{});
          else main_tile(it2);
          spin = 0;
        } else if (scr_v[qw] >= end) {
          if (hit_wr_v[qw] == rd) break;           // classified everything and nothing is pending
        } else {
          __nanosleep(32);
          if (++spin > (1u << 27)) __trap();
        }
      }
    };
    auto fused_barrier_phase = [&]() {
      // ---- end of the sample phase: publish, grid barrier, the CTAs select the queries' bounds, grid barrier ----
      const unsigned n_ctas = gridDim.x * gridDim.y;
      if (!screen_on) {                        // (with screening the screener has published its tile maxima)
        float* mine = fa.samp + ((size_t)(q0 + ql) * n_slices + slice) * SM_SAMPLE_TOP;
#pragma unroll
        for (int e = 0; e < SM_SAMPLE_TOP; ++e) mine[e] = (q_valid && top_i[e] >= 0) ? top_s[e] : -INFINITY;
      }
      __threadfence();
      ptx::named_bar_sync(2, SM_MQ);
      if (ql == 0) {
        atomicAdd(fa.counters, 1u);
        grid_wait(fa.counters, n_ctas);
        __threadfence();
      }
      ptx::named_bar_sync(2, SM_MQ);
      // CTA c selects the bounds of queries c, c + n_ctas, ... (one query per CTA when B <= #CTAs)
      for (int qq = slice * (int)gridDim.x + qt; qq < B; qq += (int)n_ctas) {
        float smax;
        const float kth = kth_largest_128(list_s, reinterpret_cast<uint32_t*>(list_i),
                                          fa.samp + (size_t)qq * n_slices * SM_SAMPLE_TOP, n_slices * SM_SAMPLE_TOP, k, ql, smax);
        if (ql == 0) {
          tau_g[qq] = kth;
          if (fa.hpar) fa.hpar[qq] = hist_params(kth, smax);
        }
        ptx::named_bar_sync(2, SM_MQ);
      }
      __threadfence();
      ptx::named_bar_sync(2, SM_MQ);
      if (ql == 0) {
        atomicAdd(fa.counters + 1, 1u);
        grid_wait(fa.counters + 1, n_ctas);
        __threadfence();
      }
      ptx::named_bar_sync(2, SM_MQ);
      tg = q_valid ? __ldcg(tau_g + q) : INFINITY;
      tg_pending = tg;
      load_hist_params();
      if (screen_on) {
        *(volatile float*)(thr_sh + ql) = tg;
        __threadfence_block();
        __syncwarp();
        if (lane == 0) *(volatile int*)&go_main[qw] = 1;
      }
    };
    if (screen_on) {
      if (FUSED) {
        drain(S, std::true_type{});
        fused_barrier_phase();
        drain(n_seq, std::false_type{});
      } else if (SAMPLE) {
        drain(n_seq, std::true_type{});
      } else {
        drain(n_seq, std::false_type{});
      }
    } else if (FUSED) {
      for (it = 0; it < S; ++it) tile_body(it, std::true_type{});
      fused_barrier_phase();
      for (; it < n_seq; ++it) main_tile(it);
    } else if (SAMPLE) {
      for (it = 0; it < n_seq; ++it) tile_body(it, std::true_type{});
    } else {
      for (it = 0; it < n_seq; ++it) main_tile(it);
    }
    if (threadIdx.x == 128) SM_MARK(2);      // last tile filtered
    if (SAMPLE) {
      if (q_valid && !screen_on) {
        int n_keep = 0;
#pragma unroll
        for (int e = 0; e < SM_SAMPLE_TOP; ++e) n_keep += top_i[e] >= 0 ? 1 : 0;
        size_t ob = (size_t)q * (size_t)cap + (size_t)atomicAdd(out_n + q, n_keep);
#pragma unroll
        for (int e = 0; e < SM_SAMPLE_TOP; ++e)
          if (top_i[e] >= 0) { out_s[ob] = top_s[e]; out_i[ob] = top_i[e]; ++ob; }
      }
    } else {
      publish();
    }
    __syncwarp();
    if (lane == 0) atomicAdd(&keepers_done, 1);          // the sweepers (warps 2-3) leave their loop at 4
  }

  if (threadIdx.x == 128) SM_MARK(3);        // candidates published
  ptx::tc_fence_before_sync();
  // pair: no CTA may leave (or free tensor memory) while the peer's MMAs / commits / remote arrives still target it
  if (PAIR) ptx::cluster_sync(); else __syncthreads();
  if (warp == 2) {
    if (PAIR) ptx::tmem_dealloc_2cta(tmem_base, SM_TMEM_COLS);
    else ptx::tmem_dealloc(tmem_base, SM_TMEM_COLS);
  }
  if (threadIdx.x == 0) SM_MARK(4);          // kernel exit
  if (trace && (dbg & (1 << 20)) && threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y < SM_TRACE_TILES) {
    long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    trace[6 * SM_TRACE_TILES + blockIdx.y] = gt;
  }
#undef SM_MARK
}

template <int MODE, bool PAIR>
__global__ void __launch_bounds__(SM_THREADS, 1)
score_topk_mma_kernel(const float* __restrict__ Q, const __grid_constant__ CUtensorMap map_d, int B, int64_t N,
                      int k, int n_slices, float* __restrict__ tau_g, float* __restrict__ out_s,
                      int32_t* __restrict__ out_i, int32_t* __restrict__ out_n, long long* __restrict__ trace,
                      int dbg, FusedArgs fa, int seg_tiles, int cap) {
  scorer_body<MODE, PAIR>(Q, map_d, B, N, k, n_slices, tau_g, out_s, out_i, out_n, trace, dbg, fa, seg_tiles, cap);
}

// ---------------------------------------------------------------------------------------------
// Select-merge: one CTA per query picks the exact top-k of the candidates the scoring CTAs
// published (~n_slices * 50..64 of them: 9 k per query at 148 slices) and sorts those k.
//
// Every candidate becomes one 64-bit key, (order-preserving score bits << 32) | ~index, so
// "better" (higher score, then lower index — the total order of every other top-k stage) is a
// plain unsigned compare and all keys are distinct.  The keys are staged in shared memory with
// many independent loads in flight, the k-th largest key is found by a most-significant-digit
// radix select (8-bit digits, shared-memory histograms, starting below the bits all keys share),
// exactly k keys are collected and a 64-slot bitonic network orders them.
// The previous merge streamed the candidates through warp buffers with a dependent global load
// per 256 candidates: ~65 us per call whatever the count (profiles/r1_search_b128_launch_list.md).
constexpr int MG_THREADS = 512;
constexpr int MG_MAX_STAGE = 20480;        // keys staged in shared memory (160 KB); larger sets are re-read from L2

__device__ __forceinline__ uint64_t cand_key(float s, int32_t ix) {
  return ((uint64_t)f2key(s + 0.0f) << 32) | (uint64_t)(0xffffffffu - (uint32_t)ix);   // -0 -> +0
}

template <bool STAGED>
__device__ __forceinline__ uint64_t mg_key(const uint64_t* keys, const float* cs, const int32_t* ci, int t) {
  if (STAGED) return keys[t];
  return cand_key(__ldcg(cs + t), __ldcg(ci + t));
}

template <bool STAGED>
__device__ __forceinline__ void select_merge_body(uint64_t* keys, const float* cs, const int32_t* ci, int total,
                                                  int k, int64_t idx_offset, float* res_s, int64_t* res_i) {
  __shared__ uint32_t hist[256];
  __shared__ uint64_t red_and[MG_THREADS / 32], red_or[MG_THREADS / 32];
  __shared__ uint64_t sel[TOPK_KMAX];
  __shared__ uint64_t s_prefix;
  __shared__ int s_remaining, s_done, s_nout, s_shift;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  uint64_t thresh = 0;                       // keys >= thresh are the answer
  if (total > k) {
    // bits shared by all keys need no histogram pass: AND/OR reduction finds them
    uint64_t a = ~0ull, o = 0ull;
    for (int t = tid; t < total; t += MG_THREADS) {
      const uint64_t key = mg_key<STAGED>(keys, cs, ci, t);
      a &= key; o |= key;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      a &= __shfl_xor_sync(0xffffffffu, a, d);
      o |= __shfl_xor_sync(0xffffffffu, o, d);
    }
    if (lane == 0) { red_and[warp] = a; red_or[warp] = o; }
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < MG_THREADS / 32; ++w) { a &= red_and[w]; o |= red_or[w]; }
      const uint64_t diff = a ^ o;           // non-zero: the keys are distinct and total > k >= 1
      const int top = 63 - __clzll((long long)diff);          // highest differing bit
      const int shift = (top >> 3) << 3;                      // first digit = the byte holding that bit
      s_shift = shift;
      s_prefix = shift + 8 >= 64 ? 0ull : (a >> (shift + 8)) << (shift + 8);
      s_remaining = k;
      s_done = 0;
    }
    __syncthreads();
    int shift = s_shift;
    while (true) {
      const uint64_t prefix = s_prefix;
      const uint64_t hi_mask = shift + 8 >= 64 ? 0ull : (~0ull << (shift + 8));
      for (int b = tid; b < 256; b += MG_THREADS) hist[b] = 0;
      __syncthreads();
      for (int t = tid; t < total; t += MG_THREADS) {
        const uint64_t key = mg_key<STAGED>(keys, cs, ci, t);
        if ((key & hi_mask) == (prefix & hi_mask)) atomicAdd(&hist[(uint32_t)(key >> shift) & 255u], 1u);
      }
      __syncthreads();
      if (warp == 0) {
        // lane l owns bins [8l, 8l+8); suffix sums over lanes find the bin holding the remaining-th key from the top
        uint32_t h[8];
        uint32_t own = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) { h[j] = hist[8 * lane + j]; own += h[j]; }
        uint32_t suf = own;                  // inclusive suffix sum over lanes >= lane
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const uint32_t v = __shfl_down_sync(0xffffffffu, suf, d);
          if (lane + d < 32) suf += v;
        }
        const uint32_t above = suf - own;    // keys in bins of higher lanes
        const uint32_t rem = (uint32_t)s_remaining;
        __syncwarp();                        // every lane has read s_remaining before one lane rewrites it
        if (above < rem && rem <= suf) {
          uint32_t cum = above;
#pragma unroll
          for (int j = 7; j >= 0; --j) {
            if (cum < rem && rem <= cum + h[j]) {
              s_prefix = (prefix & hi_mask) | ((uint64_t)(8 * lane + j) << shift);
              s_remaining = (int)(rem - cum);
              s_done = (h[j] == rem - cum) || shift == 0;      // the whole bin is taken: no need to look inside
            }
            cum += h[j];
          }
        }
      }
      __syncthreads();
      if (s_done) break;
      shift = shift >= 8 ? shift - 8 : 0;
    }
    thresh = s_prefix;                       // low bits zero: every key of the selected bin (and above) qualifies
  }
  if (tid == 0) s_nout = 0;
  for (int j = tid; j < TOPK_KMAX; j += MG_THREADS) sel[j] = 0ull;
  __syncthreads();
  for (int t = tid; t < total; t += MG_THREADS) {
    const uint64_t key = mg_key<STAGED>(keys, cs, ci, t);
    if (key >= thresh) {
      const int pos = atomicAdd(&s_nout, 1);
      if (pos < TOPK_KMAX) sel[pos] = key;
    }
  }
  __syncthreads();
  if (warp == 0) {
    // 64-slot bitonic network, descending; empty slots hold key 0 (below every real key)
    for (int kk = 2; kk <= TOPK_KMAX; kk <<= 1) {
      for (int j = kk >> 1; j > 0; j >>= 1) {
        const int i = ((lane & ~(j - 1)) << 1) | (lane & (j - 1));
        const int p = i | j;
        const bool desc = (i & kk) == 0;
        const uint64_t ki = sel[i], kp = sel[p];
        if ((kp > ki) == desc) { sel[i] = kp; sel[p] = ki; }
        __syncwarp();
      }
    }
    for (int j = lane; j < k; j += 32) {
      const uint64_t key = sel[j];
      const bool valid = key != 0ull;
      res_s[j] = valid ? key2f((uint32_t)(key >> 32)) : -INFINITY;
      res_i[j] = valid ? (int64_t)(0xffffffffu - (uint32_t)key) + idx_offset : (int64_t)-1;
    }
  }
}

__global__ void __launch_bounds__(MG_THREADS)
topk_select_merge_kernel(const float* __restrict__ out_s, const int32_t* __restrict__ out_i,
                         const int32_t* __restrict__ out_n, int cap_i, int k, int64_t idx_offset, int stage_cap,
                         float* __restrict__ res_s, int64_t* __restrict__ res_i) {
  extern __shared__ uint64_t mg_keys[];
  const int q = blockIdx.x;
  const size_t cap = (size_t)cap_i;
  const int total = out_n[q];
  const float* cs = out_s + (size_t)q * cap;
  const int32_t* ci = out_i + (size_t)q * cap;
  float* rs = res_s + (int64_t)q * k;
  int64_t* ri = res_i + (int64_t)q * k;
  if (total <= stage_cap) {
    // stage: four independent load pairs in flight per thread
    int t = threadIdx.x;
    for (; t + 3 * MG_THREADS < total; t += 4 * MG_THREADS) {
      float s[4];
      int32_t ix[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { s[u] = __ldcg(cs + t + u * MG_THREADS); ix[u] = __ldcg(ci + t + u * MG_THREADS); }
#pragma unroll
      for (int u = 0; u < 4; ++u) mg_keys[t + u * MG_THREADS] = cand_key(s[u], ix[u]);
    }
    for (; t < total; t += MG_THREADS) mg_keys[t] = cand_key(__ldcg(cs + t), __ldcg(ci + t));
    __syncthreads();
    select_merge_body<true>(mg_keys, cs, ci, total, k, idx_offset, rs, ri);
  } else {
    select_merge_body<false>(mg_keys, cs, ci, total, k, idx_offset, rs, ri);
  }
}

// stage capacity (keys) for a given per-query candidate capacity; bytes = 8 * capacity
static int merge_stage_cap(int cap) { return std::min(cap, MG_MAX_STAGE); }

static int launch_select_merge(const float* outs, const int32_t* outi, const int32_t* outn, int B, int cap, int k,
                               int64_t idx_offset, float* res_s, int64_t* res_i, cudaStream_t st) {
  const int stage_cap = (g_debug_flags & 512) ? 0 : merge_stage_cap(cap);   // bit 9: force the L2 re-read path
  const size_t smem = (size_t)stage_cap * 8;
  static thread_local int attr_dev = -1;            // once per host thread and device: five driver calls per search otherwise
  int cur_dev = 0;
  TTR_CHECK_CUDA(cudaGetDevice(&cur_dev));
  if (attr_dev != cur_dev) {
    TTR_CHECK_CUDA(cudaFuncSetAttribute(topk_select_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        MG_MAX_STAGE * 8));
    attr_dev = cur_dev;
  }
  topk_select_merge_kernel<<<B, MG_THREADS, smem, st>>>(outs, outi, outn, cap, k, idx_offset, stage_cap, res_s, res_i);
  TTR_CHECK_LAUNCH();
  return TTR_OK;
}

struct MmaPlan {
  int n_qt, n_slices, n_ctas;
  bool pair;                        // CTA pairs (cta_group::2, 64-document tiles): every batch of more than one query tile
  int nd_t;                         // documents per tile: 32, pair 64
  int seg_tiles, n_segs, cap;       // id-segment length (tiles), segments per CTA, candidate slots per query
  int64_t tau_off, outs_off, outi_off, outn_off, samp_off, cnt_off, hist_off, hpar_off, total;
};

MmaPlan mma_plan(int B, int64_t N) {
  MmaPlan p;
  const int n_qt = ceil_div(B, SM_MQ);
  p.pair = n_qt >= 2 && !(g_debug_flags & (1 << 27));      // debug bit 27: one CTA per query tile as in round 1 (A/B)
  p.n_qt = p.pair ? (n_qt + 1) / 2 * 2 : n_qt;              // grid.x: a pair needs two tiles (the odd one out has no valid query)
  p.nd_t = p.pair ? 2 * SM_ND : SM_ND;
  const int sms = sm_count();
  p.n_slices = std::max(1, sms / p.n_qt);      // one wave: n_qt * n_slices <= #SMs (1 CTA per SM)
  p.n_ctas = p.n_qt * p.n_slices;
  // a list entry numbers its document with 16 bits inside a segment of <= 65536 documents; the CTA publishes its list at
  // every segment boundary (debug bit 24: 16-tile segments, so small test corpora cross many boundaries)
  p.seg_tiles = (g_debug_flags & (1 << 24)) ? 16 : 65536 / p.nd_t;
  const int64_t n_tiles = ceil_div64(N, (int64_t)p.nd_t);
  p.n_segs = (int)std::max<int64_t>(1, ceil_div64(ceil_div64(n_tiles, (int64_t)p.n_slices), (int64_t)p.seg_tiles));
  p.cap = p.n_slices * p.n_segs * SM_KEEP;
  auto align = [](int64_t v) { return (v + 255) / 256 * 256; };
  const int64_t bp = (int64_t)p.n_qt * SM_MQ;
  p.tau_off = 0;
  p.outs_off = align(bp * 4);
  p.outi_off = align(p.outs_off + bp * p.cap * 4);
  p.outn_off = align(p.outi_off + bp * p.cap * 4);
  p.samp_off = align(p.outn_off + bp * 4);                                   // fused mode: [queries][n_slices][4] fp32
  p.cnt_off = align(p.samp_off + bp * p.n_slices * SM_SAMPLE_TOP * 4);
  p.hist_off = align(p.cnt_off + 16);                                        // [queries][SM_HBINS] u32
  p.hpar_off = align(p.hist_off + bp * SM_HBINS * 4);                        // [queries] float2
  p.total = align(p.hpar_off + bp * 8);
  return p;
}

// One launch path for the six instantiations: cluster dimension (2,1,1) for pairs, cooperative for the fused mode.
static cudaError_t launch_scorer(const void* kern, dim3 grid, size_t smem, cudaStream_t st, bool pair, bool coop, void** args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(SM_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[2];
  int n = 0;
  if (pair) {
    at[n].id = cudaLaunchAttributeClusterDimension;
    at[n].val.clusterDim.x = 2; at[n].val.clusterDim.y = 1; at[n].val.clusterDim.z = 1;
    ++n;
  }
  if (coop) {
    at[n].id = cudaLaunchAttributeCooperative;
    at[n].val.cooperative = 1;
    ++n;
  }
  cfg.attrs = at;
  cfg.numAttrs = n;
  return cudaLaunchKernelExC(&cfg, kern, args);
}

template <int MODE>
static const void* scorer_fn(bool pair) {
  return pair ? (const void*)score_topk_mma_kernel<MODE, true> : (const void*)score_topk_mma_kernel<MODE, false>;
}

int launch_score_topk_mma(const float* Q, int B, const float* docs, int64_t N, int k, int64_t row_offset,
                          void* workspace, float* out_scores, int64_t* out_idx, cudaStream_t st) {
  MmaPlan p = mma_plan(B, N);
  unsigned char* ws = reinterpret_cast<unsigned char*>(workspace);
  float* tau = reinterpret_cast<float*>(ws + p.tau_off);
  float* outs = reinterpret_cast<float*>(ws + p.outs_off);
  int32_t* outi = reinterpret_cast<int32_t*>(ws + p.outi_off);
  int32_t* outn = reinterpret_cast<int32_t*>(ws + p.outn_off);
  const int nq_pad = p.n_qt * SM_MQ;
  float* samp = reinterpret_cast<float*>(ws + p.samp_off);
  unsigned int* counters = reinterpret_cast<unsigned int*>(ws + p.cnt_off);
  unsigned int* hist = reinterpret_cast<unsigned int*>(ws + p.hist_off);
  float2* hpar = reinterpret_cast<float2*>(ws + p.hpar_off);
  init_tau_kernel<<<std::min(ceil_div(nq_pad * SM_HBINS, 256), 2 * sm_count()), 256, 0, st>>>(tau, outn, nq_pad, counters, hist, hpar);
  TTR_CHECK_LAUNCH();
  const size_t smem = (size_t)SM_STAGES * SM_STAGE_BYTES + SM_LIST_BYTES + SM_LISTI_BYTES + (3 * SM_STAGES + 2 * SM_NACC_MAX) * 8 + 16 + 1024;
  static thread_local int attr_dev = -1;
  int cur_dev = 0;
  TTR_CHECK_CUDA(cudaGetDevice(&cur_dev));
  if (attr_dev != cur_dev) {
    for (int pr = 0; pr < 2; ++pr) {
      TTR_CHECK_CUDA(cudaFuncSetAttribute(scorer_fn<0>(pr), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      TTR_CHECK_CUDA(cudaFuncSetAttribute(scorer_fn<1>(pr), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      TTR_CHECK_CUDA(cudaFuncSetAttribute(scorer_fn<2>(pr), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    attr_dev = cur_dev;
  }
  // The survivor histogram pays where a CTA's own lists cannot tighten the bound fast enough: short scans (<= 280 tiles
  // per CTA: 1/8 of the corpus and less at B <= 128) and CTA pairs (64-document tiles, twice the survivors per tile).
  // Same-box A/Bs (r2, B = 128, histogram off -> on): 0.8 M docs 0.210 -> 0.180 ms, 1.1 M 0.240 -> 0.229, 1.6 M 0.290 -> 0.301,
  // 2.2 M 0.377 -> 0.397, 4.4 M 0.714 -> 0.719, 8.84 M 1.326 -> 1.381 (lists fill by themselves, the sweeps and REDs only
  // cost); pairs: 1 M docs B = 256 0.375 -> 0.256, 8.84 M 2.05 -> 1.85.
  const int64_t scan_tiles_per_cta = ceil_div64(ceil_div64(N, (int64_t)p.nd_t), (int64_t)p.n_slices);
  // (TTR_HIST_MAX_TILES overrides the 280 for the A/B, read once)
  static const int64_t hist_max_tiles = [] { const char* e = getenv("TTR_HIST_MAX_TILES"); return e ? (int64_t)atoll(e) : (int64_t)280; }();
  const bool use_hist = p.pair || scan_tiles_per_cta <= hist_max_tiles;
  if (!use_hist) hist = nullptr;
  FusedArgs fa{nullptr, nullptr, 0, hist, hpar};
  long long* tr = nullptr;
  int dbgv = g_debug_flags;
  CUtensorMap map;
  int64_t n_scan = N;
  int n_sl = p.n_slices, seg = p.seg_tiles, cap = p.cap;
  void* args[] = {(void*)&Q, (void*)&map, (void*)&B, (void*)&n_scan, (void*)&k, (void*)&n_sl, (void*)&tau, (void*)&outs,
                  (void*)&outi, (void*)&outn, (void*)&tr, (void*)&dbgv, (void*)&fa, (void*)&seg, (void*)&cap};
  // Sample pass: exact top-k of the first ~38 k documents gives every query a k-th-best bound
  // (top ~0.1 %) before the full scan starts.  Without it each CTA spends its first ~250
  // tiles appending and compacting almost everything it sees (3-4 k cycles per tile instead
  // of ~1.2 k; profiles/r1_score_topk_mma_v4_trace_*).  The bound is valid for any document
  // order; its tightness only matters for speed.
  // Sample size: 1/20 of a CTA's tiles, between 4 and 32 per SM.  Since the sample epilogue keeps a register
  // top-4 instead of lists the pass is ~35 us fixed + 1-2 us per tile, and a tighter bound saves the main pass more
  // than that (measured per 128-query step: 8.8 M docs 8 -> 16 -> 32 tiles 1.62 -> 1.56 -> 1.50 ms; 1.1 M docs
  // 4 -> 8 -> 12 tiles 0.295 -> 0.275 -> 0.244 ms with the cheaper epilogue)
  const int tiles_override = (g_debug_flags >> 12) & 63;          // bits 12-17: sample tiles per SM (experiments)
  const int64_t tiles_per_cta = ceil_div64(ceil_div64(N, (int64_t)SM_ND), (int64_t)sm_count());
  const int tiles_auto = (int)std::min<int64_t>(32, std::max<int64_t>(4, tiles_per_cta / 20));
  const int64_t n_sample = (int64_t)sm_count() * (tiles_override ? tiles_override : tiles_auto) * SM_ND;
  // The sample publishes SM_SAMPLE_TOP scores per (CTA, query): with few slices (many query tiles) the union holds
  // fewer than ~3k values and its k-th best is no bound worth a pass -> large batches start unseeded (their CTAs
  // scan >= 8 k tiles each, the ~250-tile transient is noise there).
  const bool sample_useful = p.n_slices * SM_SAMPLE_TOP >= 3 * k;
  // all CTAs co-resident (one wave): sample phase, bound selection and main pass in ONE cooperative launch
  static int fused_ok[64][2];                 // per device and kernel flavour: 0 unknown, 1 works, -1 refused once
  // (measured: 0.263 vs 0.274 ms per 128-query step on a 1.1 M-document shard, no difference on 8.8 M documents,
  // where the two-pass path stays the default; debug bit 22 forces the fused launch there too)
  const bool fuse_size = N < 4000000 || (g_debug_flags & (1 << 22));
  if (N >= 16 * n_sample && sample_useful && !(g_debug_flags & (256 | (1 << 21))) && p.n_ctas <= sm_count() &&
      fused_ok[cur_dev & 63][p.pair] >= 0 && fuse_size) {
    int rcf = make_tf32_rowmajor_map(&map, docs, N, SM_DIM, SM_ND);
    if (rcf != TTR_OK) return rcf;
    // every query sees the same number of sample documents whatever the slice count
    fa = FusedArgs{samp, counters, (int)ceil_div64(n_sample / p.nd_t, (int64_t)p.n_slices), hist, hpar};
    tr = g_score_trace;
    cudaError_t ce = launch_scorer(scorer_fn<2>(p.pair), dim3(p.n_qt, p.n_slices), smem, st, p.pair, true, args);
    if (ce == cudaSuccess) {
      fused_ok[cur_dev & 63][p.pair] = 1;
      return launch_select_merge(outs, outi, outn, B, p.cap, k, row_offset, out_scores, out_idx, st);
    }
    (void)cudaGetLastError();                 // not co-resident on this device/partition: use the two-pass path
    fused_ok[cur_dev & 63][p.pair] = -1;
    fa = FusedArgs{nullptr, nullptr, 0, hist, hpar};
    tr = nullptr;
  }
  if (N >= 16 * n_sample && sample_useful && !(g_debug_flags & 256)) {
    MmaPlan ps = mma_plan(B, n_sample);
    int rc = make_tf32_rowmajor_map(&map, docs, n_sample, SM_DIM, SM_ND);
    if (rc != TTR_OK) return rc;
    n_scan = n_sample; n_sl = ps.n_slices; seg = ps.seg_tiles; cap = ps.cap;
    TTR_CHECK_CUDA(launch_scorer(scorer_fn<1>(ps.pair), dim3(ps.n_qt, ps.n_slices), smem, st, ps.pair, false, args));
    rc = launch_select_merge(outs, outi, outn, B, ps.cap, k, 0, out_scores, out_idx, st);
    if (rc != TTR_OK) return rc;
    seed_tau_kernel<<<ceil_div(nq_pad, 256), 256, 0, st>>>(tau, outn, out_scores, out_idx, B, nq_pad, k, hpar);
    TTR_CHECK_LAUNCH();
  }
  int rc = make_tf32_rowmajor_map(&map, docs, N, SM_DIM, SM_ND);
  if (rc != TTR_OK) return rc;
  n_scan = N; n_sl = p.n_slices; seg = p.seg_tiles; cap = p.cap;
  tr = g_score_trace;
  TTR_CHECK_CUDA(launch_scorer(scorer_fn<0>(p.pair), dim3(p.n_qt, p.n_slices), smem, st, p.pair, false, args));
  return launch_select_merge(outs, outi, outn, B, p.cap, k, row_offset, out_scores, out_idx, st);
}

int64_t score_topk_mma_workspace_bytes(int B, int64_t N) { return mma_plan(B, N).total; }

}  // namespace ttr

// diagnostic: device buffer of 5 * 256 int64 that receives the pipeline timeline of CTA (0,0), or NULL
extern "C" int ttr_debug_set_trace(long long* trace) {
  ttr::g_score_trace = trace;
  return TTR_OK;
}
