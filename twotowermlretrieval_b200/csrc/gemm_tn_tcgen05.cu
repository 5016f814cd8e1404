// gemm_tn_tcgen05.cu — weight-gradient contraction C[N1, N2] (+)= A[M, N1]^T * B[M, N2] on the
// tensor cores: both operands are "MN-major" (the token dimension M is the reduction), so the
// TMA boxes are {32 columns, 32 tokens} and the UMMA descriptors use the MN-major 128B/32B-atom
// canonical layout (atoms of 4 token rows x 128 bytes; leading byte offset = stride between
// 32-column blocks, stride byte offset = stride between 4-token groups).  Split-K over CTAs, the
// partial tiles are accumulated into C by the copy engine (cp.reduce.async.bulk.tensor .add).
//
// Replaces autograd's dW = dG^T X of nn.GRU / nn.Linear reached from `loss.backward()`
// (backend/main.py:254).  The number of valid tokens is read on the device; the caller zeroes
// the rows [m_valid, round_up(m_valid, 32)) of both operands (ttr_zero_tail_rows).
#include <cudaTypedefs.h>

#include "common.cuh"
#include "ptx.cuh"

namespace ttr {

int make_pitched_map(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t pitch, int box_rows,
                     bool tf32, bool atom32 = false);

constexpr int TN_M = 128;                 // C tile rows  (columns of A)
constexpr int TN_N = 128;                 // C tile cols  (columns of B)
constexpr int TN_K = 32;                  // tokens per stage
constexpr int TN_STAGES = 6;
constexpr int TN_BLK_BYTES = TN_K * 128;  // one {32 col, 32 token} box: 4 KB
constexpr int TN_A_BYTES = (TN_M / 32) * TN_BLK_BYTES;   // 16 KB
constexpr int TN_B_BYTES = (TN_N / 32) * TN_BLK_BYTES;   // 16 KB
constexpr int TN_STAGE_BYTES = TN_A_BYTES + TN_B_BYTES;
constexpr int TN_OUT_BYTES = TN_M * 128;  // staging of a 128 x 32 output chunk
constexpr int TN_THREADS = 256;

__device__ __forceinline__ uint64_t make_mnmajor_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  // MN-major 32-bit operands only exist in the SWIZZLE_128B_BASE32B layout (atoms of 4 token rows
  // x 128 bytes, 32-byte chunks XOR-ed with the row: Swizzle<2,5,2>), which the copy engine
  // produces with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.
  d |= (uint64_t)(TN_BLK_BYTES >> 4) << 16;      // leading byte offset: next 32-column block
  d |= (uint64_t)(512 >> 4) << 32;               // stride byte offset: next group of 4 tokens
  d |= (uint64_t)1 << 46;                        // descriptor version
  d |= (uint64_t)1 << 61;                        // SWIZZLE_128B_BASE32B
  return d;
}

__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* map, const void* smem_src, int32_t c0, int32_t c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map),
               "r"(ptx::smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}

__global__ void __launch_bounds__(TN_THREADS, 1)
gemm_tn_tf32_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                    const __grid_constant__ CUtensorMap map_c, int N1, int N2, int m_bound,
                    const int32_t* __restrict__ m_valid, int splits) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* tiles = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* out_stage = tiles + TN_STAGES * TN_STAGE_BYTES;      // [2][128 rows][128 B]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(out_stage + 2 * TN_OUT_BYTES);
  uint64_t* empty_bar = full_bar + TN_STAGES;
  uint64_t* acc_full = empty_bar + TN_STAGES;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int M = m_valid ? min(m_bound, *m_valid) : m_bound;
  const int k_blocks = ceil_div(M, TN_K);
  const int t1 = ceil_div(N1, TN_M), t2 = ceil_div(N2, TN_N);
  const int kb_per = ceil_div(k_blocks, splits);
  const int total = t1 * t2 * splits;

  if (threadIdx.x == 0) {
    for (int s = 0; s < TN_STAGES; ++s) { ptx::mbar_init(full_bar + s, 1); ptx::mbar_init(empty_bar + s, 1); }
    for (int b = 0; b < 2; ++b) { ptx::mbar_init(acc_full + b, 1); ptx::mbar_init(acc_empty + b, 4); }
    ptx::fence_mbar_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_slot, 2 * TN_N);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    int it = 0;
    for (int w = blockIdx.x; w < total; w += gridDim.x) {
      const int sp = w % splits, tile = w / splits;
      const int i0 = (tile / t2) * TN_M, j0 = (tile % t2) * TN_N;
      const int kb0 = sp * kb_per, kb1 = min(k_blocks, kb0 + kb_per);
      for (int kb = kb0; kb < kb1; ++kb, ++it) {
        const int s = it % TN_STAGES;
        const uint32_t ph = (uint32_t)(it / TN_STAGES) & 1u;
        ptx::mbar_wait(empty_bar + s, ph ^ 1u);
        unsigned char* a_dst = tiles + s * TN_STAGE_BYTES;
        if (ptx::elect_one()) {
          ptx::mbar_arrive_expect_tx(full_bar + s, TN_STAGE_BYTES);
#pragma unroll
          for (int g = 0; g < TN_M / 32; ++g)
            ptx::tma_load_2d(a_dst + g * TN_BLK_BYTES, &map_a, i0 + g * 32, kb * TN_K, full_bar + s);
#pragma unroll
          for (int g = 0; g < TN_N / 32; ++g)
            ptx::tma_load_2d(a_dst + TN_A_BYTES + g * TN_BLK_BYTES, &map_b, j0 + g * 32, kb * TN_K, full_bar + s);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    // kind::tf32, fp32 accumulate, A and B MN-major (instruction descriptor bits 15 and 16)
    constexpr uint32_t idesc = ptx::make_idesc_tf32(TN_M, TN_N) | (1u << 15) | (1u << 16);
    int it = 0, local = 0;
    for (int w = blockIdx.x; w < total; w += gridDim.x) {
      const int sp = w % splits;
      const int kb0 = sp * kb_per, kb1 = min(k_blocks, kb0 + kb_per);
      if (kb0 >= kb1) continue;                       // empty split: no accumulator is produced
      const int buf = local & 1;
      const uint32_t aph = (uint32_t)(local >> 1) & 1u;
      ++local;
      ptx::mbar_wait(acc_empty + buf, aph ^ 1u);
      ptx::tc_fence_after_sync();
      const uint32_t d_tmem = tmem_base + buf * TN_N;
      for (int kb = kb0; kb < kb1; ++kb, ++it) {
        const int s = it % TN_STAGES;
        const uint32_t ph = (uint32_t)(it / TN_STAGES) & 1u;
        ptx::mbar_wait(full_bar + s, ph);
        ptx::tc_fence_after_sync();
        const uint32_t a_addr = ptx::smem_u32(tiles + s * TN_STAGE_BYTES);
        const uint64_t a_desc = make_mnmajor_sw128_desc(a_addr);
        const uint64_t b_desc = make_mnmajor_sw128_desc(a_addr + TN_A_BYTES);
        if (ptx::elect_one()) {
#pragma unroll
          for (int k = 0; k < TN_K / 8; ++k)            // 8 tokens per MMA = one 1024-byte row group
            ptx::mma_tf32_ss(d_tmem, a_desc + (uint64_t)(k * (1024 >> 4)), b_desc + (uint64_t)(k * (1024 >> 4)), idesc,
                             (kb > kb0 || k > 0) ? 1u : 0u);
          ptx::mma_commit(empty_bar + s);
          if (kb == kb1 - 1) ptx::mma_commit(acc_full + buf);
        }
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue: TMEM -> swizzled smem staging -> TMA reduce-add into C =====
    const int q = warp - 4;
    const int r_in_tile = q * 32 + lane;
    int local = 0, chunk_no = 0;
    for (int w = blockIdx.x; w < total; w += gridDim.x) {
      const int sp = w % splits, tile = w / splits;
      const int kb0 = sp * kb_per, kb1 = min(k_blocks, kb0 + kb_per);
      if (kb0 >= kb1) continue;
      const int i0 = (tile / t2) * TN_M, j0 = (tile % t2) * TN_N;
      const int buf = local & 1;
      const uint32_t aph = (uint32_t)(local >> 1) & 1u;
      ++local;
      ptx::mbar_wait(acc_full + buf, aph);
      ptx::tc_fence_after_sync();
#pragma unroll 1
      for (int c0 = 0; c0 < TN_N; c0 += 32, ++chunk_no) {
        uint32_t r[32];
        ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * TN_N + c0, r);
        ptx::tmem_ld_wait();
        if (c0 + 32 >= TN_N) {
          ptx::tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(acc_empty + buf);
        }
        unsigned char* stg = out_stage + (chunk_no & 1) * TN_OUT_BYTES;
        if (warp == 4 && lane == 0) ptx::bulk_wait_group_read<1>();
        ptx::named_bar_sync(1, 128);
        float4* row = reinterpret_cast<float4*>(stg + r_in_tile * 128);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4 o;
          o.x = __uint_as_float(r[4 * j + 0]); o.y = __uint_as_float(r[4 * j + 1]);
          o.z = __uint_as_float(r[4 * j + 2]); o.w = __uint_as_float(r[4 * j + 3]);
          row[j ^ (r_in_tile & 7)] = o;
        }
        ptx::fence_proxy_async_smem();
        ptx::named_bar_sync(1, 128);
        if (warp == 4 && lane == 0) {
          if (j0 + c0 < N2) tma_reduce_add_2d(&map_c, stg, j0 + c0, i0);
          ptx::bulk_commit_group();
        }
      }
    }
    if (warp == 4 && lane == 0) ptx::bulk_wait_group<0>();
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc(tmem_base, 2 * TN_N);
}

// rows [m_valid, min(m_bound, round_up(m_valid, 32))) of a row-major [m_bound, ld] matrix := 0
__global__ void zero_tail_rows_kernel(float* __restrict__ A, int m_bound, const int32_t* __restrict__ m_valid, int ld) {
  const int M = min(m_bound, *m_valid);
  const int end = min(m_bound, (M + 31) / 32 * 32);
  const int64_t n = (int64_t)(end - M) * ld;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    A[(int64_t)M * ld + i] = 0.f;
}

}  // namespace ttr

extern "C" int ttr_zero_tail_rows(float* A, int m_bound, const int32_t* m_valid, int ld, void* stream) {
  using namespace ttr;
  TTR_REQUIRE(m_bound >= 1 && ld >= 1 && m_valid != nullptr, "ttr_zero_tail_rows: bad arguments");
  zero_tail_rows_kernel<<<32, 256, 0, (cudaStream_t)stream>>>(A, m_bound, m_valid, ld);
  TTR_CHECK_LAUNCH();
  return TTR_OK;
}

extern "C" int ttr_gemm_tn_tf32(const float* A, int lda, const float* Bm, int ldb, float* C, int ldc, int m_bound,
                                const int32_t* m_valid, int N1, int N2, int accumulate, void* stream) {
  using namespace ttr;
  TTR_REQUIRE(m_bound >= 1 && N1 >= 1 && N2 >= 1, "ttr_gemm_tn_tf32: bad shape");
  TTR_REQUIRE(lda % 4 == 0 && ldb % 4 == 0 && ldc % 4 == 0, "ttr_gemm_tn_tf32: leading dimensions must be multiples of 4");
  TTR_REQUIRE(((uintptr_t)A & 15) == 0 && ((uintptr_t)Bm & 15) == 0 && ((uintptr_t)C & 15) == 0,
              "ttr_gemm_tn_tf32: operands must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  CUtensorMap map_a, map_b, map_c;
  // operand views: [m_bound rows, N cols] with row pitch ld (columns beyond N read as zero / are not written)
  int rc = make_pitched_map(&map_a, A, m_bound, N1, lda, TN_K, true, true);
  if (rc != TTR_OK) return rc;
  rc = make_pitched_map(&map_b, Bm, m_bound, N2, ldb, TN_K, true, true);
  if (rc != TTR_OK) return rc;
  rc = make_pitched_map(&map_c, C, N1, N2, ldc, TN_M, false);
  if (rc != TTR_OK) return rc;
  if (!accumulate) TTR_CHECK_CUDA(cudaMemset2DAsync(C, (size_t)ldc * 4, 0, (size_t)N2 * 4, N1, st));
  const int tiles = ceil_div(N1, TN_M) * ceil_div(N2, TN_N);
  const int sms = sm_count();
  int splits = std::max(1, std::min(ceil_div(m_bound, 8 * TN_K), (2 * sms) / std::max(tiles, 1)));
  const size_t smem = (size_t)TN_STAGES * TN_STAGE_BYTES + 2 * TN_OUT_BYTES + (2 * TN_STAGES + 4) * sizeof(uint64_t) + 16 + 1024;
  TTR_CHECK_CUDA(cudaFuncSetAttribute(gemm_tn_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = std::min(tiles * splits, sms);
  gemm_tn_tf32_kernel<<<grid, TN_THREADS, smem, st>>>(map_a, map_b, map_c, N1, N2, m_bound, m_valid, splits);
  TTR_CHECK_LAUNCH();
  return TTR_OK;
}
