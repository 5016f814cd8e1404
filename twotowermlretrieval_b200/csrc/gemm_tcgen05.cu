// gemm_tcgen05.cu — GRU input projection C[M,N] = A[M,K] * W[N,K]^T + bias on the 5th-gen
// tensor cores: TMA (SWIZZLE_128B) -> shared-memory ring -> tcgen05.mma kind::tf32 with the
// fp32 accumulator in TMEM (double-buffered) -> tcgen05.ld epilogue with fused bias.
//
// Replaces the `X W_ih^T + b_ih` half of nn.GRU (backend/model.py:59-62) for every timestep
// and both directions in one launch (N = dirs*3H).  Persistent: one CTA per SM walks the
// (m, n) tile list; warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator,
// warps 4-7 = epilogue.  The number of valid rows is read from device memory so the caller
// never synchronises on the packed token count.
#include <cudaTypedefs.h>

#include "common.cuh"
#include "ptx.cuh"

namespace ttr {

constexpr int GM = 128;            // tile rows  (UMMA M)
constexpr int GN = 128;            // tile cols  (UMMA N)
constexpr int GK = 32;             // fp32 elements per k-block = one 128-byte swizzle row
constexpr int G_STAGES = 6;
constexpr int G_A_BYTES = GM * GK * 4;   // 16 KB
constexpr int G_B_BYTES = GN * GK * 4;   // 16 KB
constexpr int G_STAGE_BYTES = G_A_BYTES + G_B_BYTES;
constexpr int G_ACC_COLS = GN;           // fp32 accumulator columns per buffer
constexpr int G_TMEM_COLS = 2 * G_ACC_COLS;
constexpr int G_THREADS = 256;
constexpr int G_OUT_BYTES = GM * 32 * 4;   // staging buffer of one 128 x 32 output chunk: 16 KB (x2)

extern int g_debug_flags;

__global__ void __launch_bounds__(G_THREADS, 1)
gemm_tf32_bias_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                      const __grid_constant__ CUtensorMap map_c, const float* __restrict__ bias, int m_bound,
                      const int32_t* __restrict__ m_valid, int N, int K) {
  extern __shared__ unsigned char smem_raw[];
  // [stages][A | B]; SWIZZLE_128B atoms need 1024-byte alignment in the shared window
  unsigned char* tiles = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* out_stage = tiles + G_STAGES * G_STAGE_BYTES;      // [2][128 rows][128 B], SWIZZLE_128B
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(out_stage + 2 * G_OUT_BYTES);
  uint64_t* empty_bar = full_bar + G_STAGES;
  uint64_t* acc_full = empty_bar + G_STAGES;    // [2] MMA -> epilogue
  uint64_t* acc_empty = acc_full + 2;           // [2] epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int M = m_valid ? min(m_bound, *m_valid) : m_bound;
  const int m_tiles = ceil_div(M, GM), n_tiles = ceil_div(N, GN);
  const int total_tiles = m_tiles * n_tiles;
  const int k_blocks = ceil_div(K, GK);

  if (threadIdx.x == 0) {
    for (int s = 0; s < G_STAGES; ++s) { ptx::mbar_init(full_bar + s, 1); ptx::mbar_init(empty_bar + s, 1); }
    for (int b = 0; b < 2; ++b) { ptx::mbar_init(acc_full + b, 1); ptx::mbar_init(acc_empty + b, 4); }
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
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int m0 = (tile / n_tiles) * GM, n0 = (tile % n_tiles) * GN;
      for (int kb = 0; kb < k_blocks; ++kb, ++it) {
        const int s = it % G_STAGES;
        const uint32_t ph = (uint32_t)(it / G_STAGES) & 1u;
        ptx::mbar_wait(empty_bar + s, ph ^ 1u);
        unsigned char* a_dst = tiles + s * G_STAGE_BYTES;
        if (ptx::elect_one()) {
          ptx::mbar_arrive_expect_tx(full_bar + s, G_STAGE_BYTES);
          ptx::tma_load_2d(a_dst, &map_a, kb * GK, m0, full_bar + s);
          ptx::tma_load_2d(a_dst + G_A_BYTES, &map_w, kb * GK, n0, full_bar + s);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (whole warp runs the loop; one elected lane issues) =====
    constexpr uint32_t idesc = ptx::make_idesc_tf32(GM, GN);
    int it = 0, local = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++local) {
      const int buf = local & 1;
      const uint32_t aph = (uint32_t)(local >> 1) & 1u;
      ptx::mbar_wait(acc_empty + buf, aph ^ 1u);       // epilogue drained this accumulator
      ptx::tc_fence_after_sync();
      const uint32_t d_tmem = tmem_base + buf * G_ACC_COLS;
      for (int kb = 0; kb < k_blocks; ++kb, ++it) {
        const int s = it % G_STAGES;
        const uint32_t ph = (uint32_t)(it / G_STAGES) & 1u;
        ptx::mbar_wait(full_bar + s, ph);
        ptx::tc_fence_after_sync();
        const uint32_t a_addr = ptx::smem_u32(tiles + s * G_STAGE_BYTES);
        const uint64_t a_desc = ptx::make_kmajor_sw128_desc(a_addr);
        const uint64_t b_desc = ptx::make_kmajor_sw128_desc(a_addr + G_A_BYTES);
        if (ptx::elect_one()) {
#pragma unroll
          for (int k = 0; k < GK / 8; ++k) {
            // advance 8 tf32 = 32 bytes inside the 128-byte swizzle row: +2 in 16-byte units
            ptx::mma_tf32_ss(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
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
    const int q = warp - 4;                               // TMEM lane quadrant of this warp
    const int r_in_tile = q * 32 + lane;
    int local = 0, chunk_no = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++local) {
      const int buf = local & 1;
      const uint32_t aph = (uint32_t)(local >> 1) & 1u;
      const int m0 = (tile / n_tiles) * GM, n0 = (tile % n_tiles) * GN;
      ptx::mbar_wait(acc_full + buf, aph);
      ptx::tc_fence_after_sync();
#pragma unroll 1
      for (int c0 = 0; c0 < GN; c0 += 32, ++chunk_no) {
        uint32_t r[32];
        ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * G_ACC_COLS + c0, r);
        ptx::tmem_ld_wait();
        if (c0 + 32 >= GN) {                              // last read of this accumulator: hand it back
          ptx::tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(acc_empty + buf);
        }
        unsigned char* stg = out_stage + (chunk_no & 1) * G_OUT_BYTES;
        // the store that used this staging buffer two chunks ago must have finished reading it
        if (warp == 4 && lane == 0) ptx::bulk_wait_group_read<1>();
        ptx::named_bar_sync(1, 128);
        float4* row = reinterpret_cast<float4*>(stg + r_in_tile * 128);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int n = n0 + c0 + 4 * j;
          float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
          if (bias && n + 3 < N) bv = __ldg(reinterpret_cast<const float4*>(bias + n));
          float4 o;
          o.x = __uint_as_float(r[4 * j + 0]) + bv.x;
          o.y = __uint_as_float(r[4 * j + 1]) + bv.y;
          o.z = __uint_as_float(r[4 * j + 2]) + bv.z;
          o.w = __uint_as_float(r[4 * j + 3]) + bv.w;
          row[j ^ (r_in_tile & 7)] = o;                   // SWIZZLE_128B: 16-byte chunk index XOR (row mod 8)
        }
        ptx::fence_proxy_async_smem();
        ptx::named_bar_sync(1, 128);
        if (warp == 4 && lane == 0) {
          ptx::tma_store_2d(&map_c, stg, n0 + c0, m0);
          ptx::bulk_commit_group();
        }
      }
    }
    if (warp == 4 && lane == 0) ptx::bulk_wait_group<0>();
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc(tmem_base, G_TMEM_COLS);
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

}  // namespace ttr

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
  const size_t smem = (size_t)G_STAGES * G_STAGE_BYTES + 2 * G_OUT_BYTES + (2 * G_STAGES + 4) * sizeof(uint64_t) + 16 + 1024;
  TTR_CHECK_CUDA(cudaFuncSetAttribute(gemm_tf32_bias_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int tiles = ceil_div(m_bound, GM) * ceil_div(N, GN);
  const int grid = std::min(tiles, sm_count());
  gemm_tf32_bias_kernel<<<grid, G_THREADS, smem, (cudaStream_t)stream>>>(map_a, map_w, map_c, bias, m_bound, m_valid, N, K);
  TTR_CHECK_LAUNCH();
  return TTR_OK;
}
