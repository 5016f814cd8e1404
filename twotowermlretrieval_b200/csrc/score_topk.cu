// score_topk.cu — exact cosine top-k over the document-embedding matrix, small query batches.
//
// Replaces `torch.matmul(q, D.t())` + `torch.topk` (backend/evaluators.py:185-186) and the
// ChromaDB ANN lookup (frontend/main.py:153-156).  HBM-bound streaming kernel: every warp
// reads whole 1 KB document rows with two coalesced 128-bit loads per lane, keeps the query
// block in registers, reduces the per-lane partial dot products with a halving butterfly
// (values are scattered across lanes instead of all-reduced, 5 shuffle rounds in total),
// and filters scores against a per-(warp,query) running k-th best held in shared memory.
// The [B, N] score matrix is never written.  Algorithmic traffic: N * 1024 B per pass.
// (Batches > 4 go to the tcgen05 kernel in score_topk_mma.cu; debug flag bit 2 forces this
// kernel for any batch so the tests can compare the two paths.)
#include "ptx.cuh"
#include "topk_common.cuh"

namespace ttr {

constexpr int DIM = 256;
constexpr int SC_WARPS = 8;
constexpr int SC_THREADS = SC_WARPS * 32;

extern int g_debug_flags;

// Halving butterfly: NV per-lane partial values -> each value fully summed in exactly
// max(NV/32,1) slots: value index = lane * (NV/32) + i for NV >= 32, lane >> (5 - log2 NV)
// otherwise (copies in the low lanes bits).
template <int NV>
__device__ __forceinline__ void reduce_scatter(float (&v)[NV], int lane) {
  int n = NV;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    if (n > 1) {
      const int h = n >> 1;
      const bool upper = (lane & o) != 0;
#pragma unroll
      for (int i = 0; i < NV / 2; ++i) {
        if (i < h) {
          float send = upper ? v[i] : v[i + h];
          float keep = upper ? v[i + h] : v[i];
          v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
      }
      n = h;
    } else {
      v[0] += __shfl_xor_sync(0xffffffffu, v[0], o);
    }
  }
}

template <int N>
struct Log2 { static constexpr int value = 1 + Log2<N / 2>::value; };
template <>
struct Log2<1> { static constexpr int value = 0; };

// QB queries per pass.  One producer warp streams 32-document (32 KB, contiguous) chunks
// into a shared-memory ring with single cp.async.bulk copies (TMA engine, mbarrier
// complete_tx); 8 consumer warps take 4 documents each per chunk.  NV = QB*4 partial values
// per lane, value index = qb * SC_DPI + d.
constexpr int SC_DPI = 4;
constexpr int SC_CHUNK = SC_WARPS * SC_DPI;        // documents per stage
constexpr int SC_STAGE_BYTES = SC_CHUNK * DIM * 4; // 32 KB

template <int QB>
__global__ void __launch_bounds__(SC_THREADS + 32, 1)
score_topk_stream_kernel(const float* __restrict__ Q, int nq, const float* __restrict__ docs, int64_t N, int k,
                         int q_base, int grid_parts, int stages, float* __restrict__ part_s,
                         int32_t* __restrict__ part_i) {
  constexpr int DPI = SC_DPI;
  constexpr int NV = QB * DPI;
  constexpr int PER_LANE = NV >= 32 ? NV / 32 : 1;
  constexpr int COPIES = NV >= 32 ? 1 : 32 / NV;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // layout: ring [stages][32 KB] | full[stages] empty[stages] | scores [W][QB][CAP] | idx | cnt | merge area
  unsigned char* ring = smem_raw;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ring + (size_t)stages * SC_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + stages;
  float* buf_s = reinterpret_cast<float*>(empty_bar + stages);
  int32_t* buf_i = reinterpret_cast<int32_t*>(buf_s + SC_WARPS * QB * TOPK_CAP);
  int32_t* cnts = buf_i + SC_WARPS * QB * TOPK_CAP;
  float* mrg_s = reinterpret_cast<float*>(cnts + SC_WARPS * QB);
  int32_t* mrg_i = reinterpret_cast<int32_t*>(mrg_s + SC_WARPS * TOPK_KMAX);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      ptx::mbar_init(full_bar + s, 1);
      ptx::mbar_init(empty_bar + s, SC_WARPS);
    }
    ptx::fence_mbar_init();
  }
  if (threadIdx.x < SC_WARPS * QB) cnts[threadIdx.x] = 0;
  __syncthreads();

  const int64_t n_chunks = ceil_div64(N, (int64_t)SC_CHUNK);

  if (warp == SC_WARPS) {
    // ===== producer warp: one lane drives the TMA engine =====
    if (lane == 0) {
      int it = 0;
      for (int64_t c = blockIdx.x; c < n_chunks; c += gridDim.x, ++it) {
        const int s = it % stages;
        const uint32_t ph = (uint32_t)(it / stages) & 1u;
        ptx::mbar_wait(empty_bar + s, ph ^ 1u);
        const int64_t d0 = c * SC_CHUNK;
        const int64_t nd = (N - d0 < SC_CHUNK) ? (N - d0) : SC_CHUNK;
        const uint32_t bytes = (uint32_t)(nd * DIM * 4);
        ptx::mbar_arrive_expect_tx(full_bar + s, bytes);
        ptx::bulk_g2s(ring + (size_t)s * SC_STAGE_BYTES, docs + d0 * DIM, bytes, full_bar + s);
      }
    }
  } else {
    // ===== consumer warps =====
    float* my_s = buf_s + warp * QB * TOPK_CAP;
    int32_t* my_i = buf_i + warp * QB * TOPK_CAP;
    int32_t* my_cnt = cnts + warp * QB;

    // query block in registers: this lane's 8 columns of each query
    float qr[QB][8];
#pragma unroll
    for (int qb = 0; qb < QB; ++qb) {
      if (qb < nq) {
        const float4* q4 = reinterpret_cast<const float4*>(Q + (int64_t)qb * DIM);
        float4 a = __ldg(q4 + lane), b = __ldg(q4 + 32 + lane);
        qr[qb][0] = a.x; qr[qb][1] = a.y; qr[qb][2] = a.z; qr[qb][3] = a.w;
        qr[qb][4] = b.x; qr[qb][5] = b.y; qr[qb][6] = b.z; qr[qb][7] = b.w;
      } else {
#pragma unroll
        for (int c = 0; c < 8; ++c) qr[qb][c] = 0.f;
      }
    }
    // which (query, doc-in-iteration) pairs this lane owns after the butterfly
    int own_qb[PER_LANE], own_d[PER_LANE];
#pragma unroll
    for (int i = 0; i < PER_LANE; ++i) {
      int vidx = NV >= 32 ? lane * PER_LANE + i : (lane >> (5 - Log2<NV>::value));
      own_qb[i] = vidx / DPI;
      own_d[i] = vidx % DPI;
    }
    const bool lane_acts = (lane % COPIES) == 0;
    const int my_qb = own_qb[0];   // PER_LANE == 1 for every instantiation (NV <= 32)
    float tau_s = -INFINITY;
    int32_t tau_i = IDX_PAD;

    int it = 0;
    for (int64_t c = blockIdx.x; c < n_chunks; c += gridDim.x, ++it) {
      const int s = it % stages;
      const uint32_t ph = (uint32_t)(it / stages) & 1u;
      ptx::mbar_wait(full_bar + s, ph);
      const int64_t d0 = c * SC_CHUNK + (int64_t)warp * DPI;
      const float4* st4 = reinterpret_cast<const float4*>(ring + (size_t)s * SC_STAGE_BYTES) + warp * DPI * (DIM / 4);
      float v[NV];
      {
        float4 da[DPI], db[DPI];
#pragma unroll
        for (int d = 0; d < DPI; ++d) {
          da[d] = st4[d * (DIM / 4) + lane];
          db[d] = st4[d * (DIM / 4) + 32 + lane];
        }
#pragma unroll
        for (int qb = 0; qb < QB; ++qb) {
#pragma unroll
          for (int d = 0; d < DPI; ++d) {
            float a = da[d].x * qr[qb][0];
            a = fmaf(da[d].y, qr[qb][1], a);
            a = fmaf(da[d].z, qr[qb][2], a);
            a = fmaf(da[d].w, qr[qb][3], a);
            a = fmaf(db[d].x, qr[qb][4], a);
            a = fmaf(db[d].y, qr[qb][5], a);
            a = fmaf(db[d].z, qr[qb][6], a);
            a = fmaf(db[d].w, qr[qb][7], a);
            v[qb * DPI + d] = a;
          }
        }
      }
      // the stage is in registers: hand the slot back to the producer
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(empty_bar + s);
      reduce_scatter<NV>(v, lane);

      bool any_pass = false;
#pragma unroll
      for (int i = 0; i < PER_LANE; ++i) {
        const int64_t doc = d0 + own_d[i];
        const int32_t di = (int32_t)doc;
        const bool pass = lane_acts && own_qb[i] < nq && doc < N && key_better<int32_t>(v[i], di, tau_s, tau_i);
        if (pass) {
          int pos = atomicAdd(&my_cnt[own_qb[i]], 1);   // pos < CAP guaranteed by the compaction rule
          my_s[own_qb[i] * TOPK_CAP + pos] = v[i];
          my_i[own_qb[i] * TOPK_CAP + pos] = di;
        }
        any_pass |= pass;
      }
      if (__any_sync(0xffffffffu, any_pass)) {
        __syncwarp();
        // compact every query whose buffer could overflow in the next iteration
#pragma unroll 1
        for (int qb = 0; qb < QB; ++qb) {
          int cnt = my_cnt[qb];
          if (cnt > TOPK_CAP - DPI) {
            float ts; int32_t ti;
            int kept = warp_compact<int32_t>(my_s + qb * TOPK_CAP, my_i + qb * TOPK_CAP, cnt, k, lane, ts, ti);
            __syncwarp();
            if (lane == 0) my_cnt[qb] = kept;
            if (qb == my_qb) { tau_s = ts; tau_i = ti; }
          }
        }
        __syncwarp();
      }
    }

    // final per-warp compaction
#pragma unroll 1
    for (int qb = 0; qb < nq; ++qb) {
      float ts; int32_t ti;
      int cnt = my_cnt[qb];
      __syncwarp();
      int kept = warp_compact<int32_t>(my_s + qb * TOPK_CAP, my_i + qb * TOPK_CAP, cnt, k, lane, ts, ti);
      __syncwarp();
      if (lane == 0) my_cnt[qb] = kept;
    }
  }
  __syncthreads();
  // CTA merge of the SC_WARPS sorted lists of each query (all 9 warps take part)
#pragma unroll 1
  for (int qb = 0; qb < nq; ++qb) {
    for (int t = threadIdx.x; t < SC_WARPS * TOPK_KMAX; t += blockDim.x) {
      int w = t / TOPK_KMAX, j = t % TOPK_KMAX;
      bool valid = j < cnts[w * QB + qb];
      mrg_s[t] = valid ? buf_s[(w * QB + qb) * TOPK_CAP + j] : -INFINITY;
      mrg_i[t] = valid ? buf_i[(w * QB + qb) * TOPK_CAP + j] : IDX_PAD;
    }
    __syncthreads();
    block_bitonic_desc<int32_t>(mrg_s, mrg_i, SC_WARPS * TOPK_KMAX);
    const int64_t base = ((int64_t)(q_base + qb) * grid_parts + blockIdx.x) * k;
    for (int j = threadIdx.x; j < k; j += blockDim.x) {
      part_s[base + j] = mrg_s[j];
      part_i[base + j] = mrg_i[j];
    }
    __syncthreads();
  }
}

// One CTA per query: stream P*kin candidates through warp buffers, merge, emit sorted top-k.
template <typename IdxT>
__global__ void __launch_bounds__(SC_THREADS)
topk_merge_kernel(const float* __restrict__ cand_s, const IdxT* __restrict__ cand_i, int P, int B, int kin,
                  int64_t part_stride, int64_t query_stride, int k, int64_t idx_offset,
                  float* __restrict__ out_s, int64_t* __restrict__ out_i) {
  __shared__ float buf_s[SC_WARPS][TOPK_CAP];
  __shared__ IdxT buf_i[SC_WARPS][TOPK_CAP];
  __shared__ int cnts[SC_WARPS];
  __shared__ float mrg_s[SC_WARPS * TOPK_KMAX];
  __shared__ IdxT mrg_i[SC_WARPS * TOPK_KMAX];
  const int q = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float tau_s = -INFINITY;
  IdxT tau_i = IdxTraits<IdxT>::pad();
  int cnt = 0;   // warp-uniform
  const int64_t total = (int64_t)P * kin;
  // Four rounds of candidates are fetched together (eight independent loads per thread in flight): with one round per
  // iteration a single query's 148 x 50 candidates cost 29 dependent L2 round trips (~15 us of the B = 1 call).
  constexpr int PF = 4;
  for (int64_t tb = (int64_t)warp * 32; tb < total; tb += (int64_t)PF * SC_THREADS) {
    float pre_s[PF];
    IdxT pre_i[PF];
#pragma unroll
    for (int u = 0; u < PF; ++u) {
      const int64_t t = tb + (int64_t)u * SC_THREADS + lane;
      pre_s[u] = -INFINITY;
      pre_i[u] = IdxTraits<IdxT>::pad();
      if (t < total) {
        const int64_t p = t / kin, j = t % kin;
        pre_s[u] = cand_s[p * part_stride + (int64_t)q * query_stride + j];
        pre_i[u] = cand_i[p * part_stride + (int64_t)q * query_stride + j];
      }
    }
#pragma unroll
    for (int u = 0; u < PF; ++u) {
      if (tb + (int64_t)u * SC_THREADS >= total) break;          // warp-uniform
      const float s = pre_s[u];
      const IdxT ix = pre_i[u];
      const bool pass = key_better<IdxT>(s, ix, tau_s, tau_i) && ix != IdxTraits<IdxT>::pad();
      const unsigned m = __ballot_sync(0xffffffffu, pass);
      if (m) {
        if (pass) {
          int pos = cnt + __popc(m & ((1u << lane) - 1));
          buf_s[warp][pos] = s;
          buf_i[warp][pos] = ix;
        }
        cnt += __popc(m);
        __syncwarp();
        if (cnt > TOPK_CAP - 32) {
          cnt = warp_compact<IdxT>(buf_s[warp], buf_i[warp], cnt, k, lane, tau_s, tau_i);
          __syncwarp();
        }
      }
    }
  }
  cnt = warp_compact<IdxT>(buf_s[warp], buf_i[warp], cnt, k, lane, tau_s, tau_i);
  if (lane == 0) cnts[warp] = cnt;
  __syncthreads();
  for (int t = threadIdx.x; t < SC_WARPS * TOPK_KMAX; t += blockDim.x) {
    int w = t / TOPK_KMAX, j = t % TOPK_KMAX;
    bool valid = j < cnts[w];
    mrg_s[t] = valid ? buf_s[w][j] : -INFINITY;
    mrg_i[t] = valid ? buf_i[w][j] : IdxTraits<IdxT>::pad();
  }
  __syncthreads();
  block_bitonic_desc<IdxT>(mrg_s, mrg_i, SC_WARPS * TOPK_KMAX);
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    const bool valid = mrg_i[j] != IdxTraits<IdxT>::pad();
    out_s[(int64_t)q * k + j] = mrg_s[j];
    out_i[(int64_t)q * k + j] = valid ? (int64_t)mrg_i[j] + idx_offset : (int64_t)-1;
  }
}

// Cross-rank merge over PEER MEMORY: every rank's [B, kin] candidate lists (scores, global ids,
// optional TF-IDF payload) live in symmetric buffers mapped into this process; the kernel reads
// them through NVLink directly — the exchange and the merge are one launch, no all-gather.
// Ordering: (score desc, source position asc); shards are ascending row ranges and each list
// is sorted by (score desc, id asc), so source order == id order among equal scores.
constexpr int MAX_PEERS = 16;
struct PeerLists {
  const float* s[MAX_PEERS];
  const int64_t* i[MAX_PEERS];
  const double* t[MAX_PEERS];
};

__device__ __forceinline__ void merge_peers_body(const PeerLists& pl, int P, int kin, int k, float* __restrict__ out_s,
                                                 int64_t* __restrict__ out_i, double* __restrict__ out_t) {
  __shared__ float buf_s[SC_WARPS][TOPK_CAP];
  __shared__ int32_t buf_i[SC_WARPS][TOPK_CAP];
  __shared__ int cnts[SC_WARPS];
  __shared__ float mrg_s[SC_WARPS * TOPK_KMAX];
  __shared__ int32_t mrg_i[SC_WARPS * TOPK_KMAX];
  const int q = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float tau_s = -INFINITY;
  int32_t tau_i = IDX_PAD;
  int cnt = 0;
  const int total = P * kin;
  // Every candidate of the query is fetched ONCE, all loads of a thread in flight together (ids and scores are
  // independent loads, up to MG_U of each per thread), and the ids are kept in shared memory for the output phase.
  // The first version read id -> (if valid) score per 256 candidates and the ids again at the end: five dependent
  // round trips over NVLink (~2-3 us each) per query at 8 ranks; now it is one.
  constexpr int MG_U = MAX_PEERS * TOPK_KMAX / SC_THREADS;
  __shared__ int64_t cand_id[MAX_PEERS * TOPK_KMAX];
  float pre_s[MG_U];
  int64_t pre_i[MG_U];
#pragma unroll
  for (int u = 0; u < MG_U; ++u) {
    const int t = warp * 32 + u * SC_THREADS + lane;
    pre_s[u] = -INFINITY;
    pre_i[u] = -1;
    if (t < total) {
      const int p = t / kin, j = t % kin;
      pre_i[u] = __ldcg(pl.i[p] + (int64_t)q * kin + j);
      pre_s[u] = __ldcg(pl.s[p] + (int64_t)q * kin + j);
    }
  }
#pragma unroll
  for (int u = 0; u < MG_U; ++u) {
    const int t0 = warp * 32 + u * SC_THREADS;
    if (t0 >= total) break;                                   // warp-uniform
    const int t = t0 + lane;
    float s = -INFINITY;
    int32_t src = IDX_PAD;
    if (t < total) {
      cand_id[t] = pre_i[u];
      if (pre_i[u] >= 0) {
        s = pre_s[u];
        src = t;
      }
    }
    const bool pass = src != IDX_PAD && key_better<int32_t>(s, src, tau_s, tau_i);
    const unsigned m = __ballot_sync(0xffffffffu, pass);
    if (m) {
      if (pass) {
        int pos = cnt + __popc(m & ((1u << lane) - 1));
        buf_s[warp][pos] = s;
        buf_i[warp][pos] = src;
      }
      cnt += __popc(m);
      __syncwarp();
      if (cnt > TOPK_CAP - 32) {
        cnt = warp_compact<int32_t>(buf_s[warp], buf_i[warp], cnt, k, lane, tau_s, tau_i);
        __syncwarp();
      }
    }
  }
  cnt = warp_compact<int32_t>(buf_s[warp], buf_i[warp], cnt, k, lane, tau_s, tau_i);
  if (lane == 0) cnts[warp] = cnt;
  __syncthreads();
  for (int t = threadIdx.x; t < SC_WARPS * TOPK_KMAX; t += blockDim.x) {
    int w = t / TOPK_KMAX, j = t % TOPK_KMAX;
    bool valid = j < cnts[w];
    mrg_s[t] = valid ? buf_s[w][j] : -INFINITY;
    mrg_i[t] = valid ? buf_i[w][j] : IDX_PAD;
  }
  __syncthreads();
  block_bitonic_desc<int32_t>(mrg_s, mrg_i, SC_WARPS * TOPK_KMAX);
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    const int src = mrg_i[j];
    const bool valid = src != IDX_PAD;
    const int p = valid ? src / kin : 0, e = valid ? src % kin : 0;
    out_s[(int64_t)q * k + j] = mrg_s[j];
    out_i[(int64_t)q * k + j] = valid ? cand_id[src] : (int64_t)-1;
    if (out_t) out_t[(int64_t)q * k + j] = (valid && pl.t[p]) ? __ldcg(pl.t[p] + (int64_t)q * kin + e) : 0.0;
  }
}

__global__ void __launch_bounds__(SC_THREADS)
topk_merge_peers_kernel(PeerLists pl, int P, int kin, int k, float* __restrict__ out_s, int64_t* __restrict__ out_i,
                        double* __restrict__ out_t) {
  merge_peers_body(pl, P, kin, k, out_s, out_i, out_t);
}

// The same merge with the cross-rank barrier INSIDE the kernel: rank r tells every peer "my lists of step s are
// complete" by a release store of s into slot r of the peer's flag array (symmetric memory), then waits until its own
// array shows step s from every peer (acquire loads), then reads the peers' lists in place.  The lists were written by
// the previous kernels of this stream, i.e. they are complete in this GPU's memory before the flag leaves it.  One
// launch replaces the host-side symmetric-memory barrier (a kernel of its own) + the merge launch: ~25 us per search
// step at 8 GPUs (profiles/r2_*).  Every CTA signals (idempotent: the value is the step number) so that progress does
// not depend on which CTA is scheduled first; waits are bounded (trap instead of a hung GPU).
struct PeerFlags {
  uint32_t* f[MAX_PEERS];
};

__global__ void __launch_bounds__(SC_THREADS)
topk_exchange_merge_kernel(PeerLists pl, PeerFlags pf, int P, int my_rank, uint32_t step, int kin, int k,
                           float* __restrict__ out_s, int64_t* __restrict__ out_i, double* __restrict__ out_t) {
  if (threadIdx.x < P) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(pf.f[threadIdx.x] + my_rank), "r"(step) : "memory");
    const uint32_t* mine = pf.f[my_rank] + threadIdx.x;
    for (uint32_t spin = 0;; ++spin) {
      uint32_t v;
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
      if ((int32_t)(v - step) >= 0) break;
      __nanosleep(32);
      if (spin > (1u << 27)) __trap();
    }
  }
  __syncthreads();
  merge_peers_body(pl, P, kin, k, out_s, out_i, out_t);
}

static int simt_grid_parts() { return sm_count(); }

template <int QB>
static int launch_stream(const float* Q, int nq, const float* docs, int64_t N, int k, int q_base, int parts,
                         float* part_s, int32_t* part_i, cudaStream_t st) {
  const size_t fixed = (size_t)SC_WARPS * QB * TOPK_CAP * 8 + SC_WARPS * QB * 4 + SC_WARPS * TOPK_KMAX * 8;
  int stages = (int)((220 * 1024 - fixed - 256) / SC_STAGE_BYTES);
  if (stages > 6) stages = 6;
  const size_t smem = (size_t)stages * SC_STAGE_BYTES + 2 * stages * sizeof(uint64_t) + fixed;
  auto kern = score_topk_stream_kernel<QB>;
  TTR_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<parts, SC_THREADS + 32, smem, st>>>(Q, nq, docs, N, k, q_base, parts, stages, part_s, part_i);
  TTR_CHECK_LAUNCH();
  return TTR_OK;
}

int launch_score_topk_mma(const float* Q, int B, const float* docs, int64_t N, int k, int64_t row_offset,
                          void* workspace, float* out_scores, int64_t* out_idx, cudaStream_t st);
int64_t score_topk_mma_workspace_bytes(int B, int64_t N);

// Query batches up to this size run the CUDA-core streaming kernel (exact fp32 products,
// HBM-bound up to 4 queries per pass; 8 queries are fp32-FMA-bound at 2.5 ms vs 1.5 ms for the
// tensor-core path); larger batches run the tcgen05 kernel (128 queries per pass).
constexpr int STREAM_MAX_B = 4;

}  // namespace ttr

extern "C" int64_t ttr_score_topk_workspace_bytes(int B, int64_t N, int k) {
  if (B > ttr::STREAM_MAX_B && !(ttr::g_debug_flags & 4)) return ttr::score_topk_mma_workspace_bytes(B, N);
  int parts = ttr::simt_grid_parts();
  int64_t bp = (int64_t)((B + 7) / 8) * 8;
  return bp * parts * (int64_t)k * 8 + 256;
}

extern "C" int ttr_score_topk(const float* Q, int B, const float* docs, int64_t N, int D, int k,
                              int64_t row_offset, float* out_scores, int64_t* out_idx, void* workspace,
                              int64_t workspace_bytes, void* stream) {
  using namespace ttr;
  TTR_REQUIRE(D == DIM, "ttr_score_topk: D=%d, the fused kernels need D == 256", D);
  TTR_REQUIRE(k >= 1 && k <= TOPK_KMAX, "ttr_score_topk: k=%d outside [1, %d]", k, TOPK_KMAX);
  TTR_REQUIRE(B >= 1 && N >= 1, "ttr_score_topk: empty problem (B=%d, N=%lld)", B, (long long)N);
  TTR_REQUIRE(N < ((int64_t)1 << 31) - 1, "ttr_score_topk: shard of %lld rows exceeds int32 local ids", (long long)N);
  TTR_REQUIRE(workspace_bytes >= ttr_score_topk_workspace_bytes(B, N, k), "ttr_score_topk: workspace too small");
  TTR_REQUIRE(((uintptr_t)docs & 15) == 0 && ((uintptr_t)Q & 15) == 0, "ttr_score_topk: Q/docs must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  if (B > STREAM_MAX_B && !(g_debug_flags & 4))
    return launch_score_topk_mma(Q, B, docs, N, k, row_offset, workspace, out_scores, out_idx, st);
  const int parts = simt_grid_parts();
  const int64_t bp = (int64_t)((B + 7) / 8) * 8;
  float* part_s = reinterpret_cast<float*>(workspace);
  int32_t* part_i = reinterpret_cast<int32_t*>(part_s + bp * parts * k);
  int q0 = 0;
  while (q0 < B) {
    int rem = B - q0;
    int rc;
    if (rem >= 8) { rc = launch_stream<8>(Q + (int64_t)q0 * DIM, 8, docs, N, k, q0, parts, part_s, part_i, st); q0 += 8; }
    else if (rem >= 4) { rc = launch_stream<4>(Q + (int64_t)q0 * DIM, 4, docs, N, k, q0, parts, part_s, part_i, st); q0 += 4; }
    else if (rem >= 2) { rc = launch_stream<2>(Q + (int64_t)q0 * DIM, 2, docs, N, k, q0, parts, part_s, part_i, st); q0 += 2; }
    else { rc = launch_stream<1>(Q + (int64_t)q0 * DIM, 1, docs, N, k, q0, parts, part_s, part_i, st); q0 += 1; }
    if (rc != TTR_OK) return rc;
  }
  // partial layout [q][parts][k]: part_stride = k, query_stride = parts*k
  topk_merge_kernel<int32_t><<<B, SC_THREADS, 0, st>>>(part_s, part_i, parts, B, k, (int64_t)k,
                                                     (int64_t)parts * k, k, row_offset, out_scores, out_idx);
  TTR_CHECK_LAUNCH();
  return TTR_OK;
}

extern "C" int ttr_topk_merge(const float* cand_scores, const int64_t* cand_idx, int P, int B, int kin, int k,
                              float* out_scores, int64_t* out_idx, void* stream) {
  using namespace ttr;
  TTR_REQUIRE(k >= 1 && k <= TOPK_KMAX, "ttr_topk_merge: k=%d outside [1, %d]", k, TOPK_KMAX);
  TTR_REQUIRE(P >= 1 && B >= 1 && kin >= 1, "ttr_topk_merge: empty problem");
  // candidate layout [P][B][kin]
  topk_merge_kernel<int64_t><<<B, SC_THREADS, 0, (cudaStream_t)stream>>>(
      cand_scores, cand_idx, P, B, kin, (int64_t)B * kin, (int64_t)kin, k, 0, out_scores, out_idx);
  TTR_CHECK_LAUNCH();
  return TTR_OK;
}

extern "C" int ttr_topk_merge_peers(const uint64_t* peer_scores_h, const uint64_t* peer_idx_h,
                                    const uint64_t* peer_tfidf_h, int P, int B, int kin, int k, float* out_scores,
                                    int64_t* out_idx, double* out_tfidf, void* stream) {
  using namespace ttr;
  TTR_REQUIRE(k >= 1 && k <= TOPK_KMAX, "ttr_topk_merge_peers: k=%d outside [1, %d]", k, TOPK_KMAX);
  TTR_REQUIRE(P >= 1 && P <= MAX_PEERS && B >= 1 && kin >= 1 && kin <= TOPK_KMAX, "ttr_topk_merge_peers: bad shape (P=%d, kin=%d)", P, kin);
  PeerLists pl;
  for (int p = 0; p < MAX_PEERS; ++p) {
    pl.s[p] = p < P ? reinterpret_cast<const float*>(peer_scores_h[p]) : nullptr;
    pl.i[p] = p < P ? reinterpret_cast<const int64_t*>(peer_idx_h[p]) : nullptr;
    pl.t[p] = (p < P && peer_tfidf_h) ? reinterpret_cast<const double*>(peer_tfidf_h[p]) : nullptr;
  }
  topk_merge_peers_kernel<<<B, SC_THREADS, 0, (cudaStream_t)stream>>>(pl, P, kin, k, out_scores, out_idx, out_tfidf);
  TTR_CHECK_LAUNCH();
  return TTR_OK;
}

extern "C" int ttr_topk_exchange_merge(const uint64_t* peer_scores_h, const uint64_t* peer_idx_h,
                                       const uint64_t* peer_tfidf_h, const uint64_t* peer_flags_h, int P, int my_rank,
                                       uint32_t step, int B, int kin, int k, float* out_scores, int64_t* out_idx,
                                       double* out_tfidf, void* stream) {
  using namespace ttr;
  TTR_REQUIRE(k >= 1 && k <= TOPK_KMAX, "ttr_topk_exchange_merge: k=%d outside [1, %d]", k, TOPK_KMAX);
  TTR_REQUIRE(P >= 1 && P <= MAX_PEERS && B >= 1 && kin >= 1 && kin <= TOPK_KMAX, "ttr_topk_exchange_merge: bad shape (P=%d, kin=%d)", P, kin);
  TTR_REQUIRE(my_rank >= 0 && my_rank < P, "ttr_topk_exchange_merge: rank %d outside [0, %d)", my_rank, P);
  PeerLists pl;
  PeerFlags pf;
  for (int p = 0; p < MAX_PEERS; ++p) {
    pl.s[p] = p < P ? reinterpret_cast<const float*>(peer_scores_h[p]) : nullptr;
    pl.i[p] = p < P ? reinterpret_cast<const int64_t*>(peer_idx_h[p]) : nullptr;
    pl.t[p] = (p < P && peer_tfidf_h) ? reinterpret_cast<const double*>(peer_tfidf_h[p]) : nullptr;
    pf.f[p] = p < P ? reinterpret_cast<uint32_t*>(peer_flags_h[p]) : nullptr;
  }
  topk_exchange_merge_kernel<<<B, SC_THREADS, 0, (cudaStream_t)stream>>>(pl, pf, P, my_rank, step, kin, k, out_scores,
                                                                       out_idx, out_tfidf);
  TTR_CHECK_LAUNCH();
  return TTR_OK;
}
