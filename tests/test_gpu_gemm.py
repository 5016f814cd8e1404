"""GPU parity: tcgen05 tf32 input-projection GEMM vs fp64, and vs the fp32 CUDA-core kernel."""
import numpy as np
import pytest
import torch

from twotowermlretrieval_b200 import _lib

pytestmark = pytest.mark.gpu


def _run(name, A, W, bias, m_valid=None):
    M, K = A.shape
    N = W.shape[0]
    C = torch.full((M, N), float("nan"), device=A.device)
    mv = None if m_valid is None else torch.tensor([m_valid], dtype=torch.int32, device=A.device)
    _lib.call(name, A, W, bias, C, M, mv, N, K)
    torch.cuda.synchronize()
    return C


@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (128, 128, 64), (1, 8, 4), (300, 1536, 200), (1000, 1536, 512),
                                   (777, 96, 12), (129, 100, 36), (4096, 1536, 200)])
def test_tf32_gemm_matches_fp64(cuda_device, M, N, K):
    g = torch.Generator(device=cuda_device).manual_seed(M + N + K)
    A = torch.randn(M, K, device=cuda_device, generator=g) * 0.4
    W = (torch.rand(N, K, device=cuda_device, generator=g) - 0.5) / 8
    b = torch.randn(N, device=cuda_device, generator=g) * 0.05
    ref = (A.double() @ W.double().t() + b.double())
    C = _run("ttr_gemm_tf32_bias", A, W, b)
    scale = float((A.double().abs() @ W.double().abs().t()).max())
    err = float((C.double() - ref).abs().max())
    assert err <= 1.5e-3 * scale, (err, scale)          # tf32: 10-bit mantissa operands, fp32 accumulate
    Cf = _run("ttr_debug_gemm_fp32_bias", A, W, b)
    assert float((Cf.double() - ref).abs().max()) <= 1e-5 * max(scale, 1.0)


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (1, 8, 8), (300, 1536, 200), (1000, 1536, 512), (777, 96, 24),
                                   (129, 104, 40), (4096, 1536, 200)])
def test_f16_gemm_matches_fp64_of_the_rounded_operands(cuda_device, M, N, K):
    """kind::f16 variant (fp16 A/W/C storage, fp32 accumulation + bias): exact up to the fp32 accumulation
    order and the final fp16 rounding when compared on the SAME fp16-rounded operands."""
    g = torch.Generator(device=cuda_device).manual_seed(M + N + K)
    A = (torch.randn(M, K, device=cuda_device, generator=g) * 0.4).half()
    W32 = (torch.rand(N, K, device=cuda_device, generator=g) - 0.5) / 8
    b = torch.randn(N, device=cuda_device, generator=g) * 0.05
    W = torch.empty(N, K, dtype=torch.float16, device=cuda_device)
    _lib.call("ttr_f32_to_f16", W32, W, W32.numel())
    assert torch.equal(W, W32.half())
    C = torch.full((M, N), float("nan"), dtype=torch.float16, device=cuda_device)
    _lib.call("ttr_gemm_f16_bias", A, W, b, C, M, None, N, K)
    torch.cuda.synchronize()
    ref = A.double() @ W.double().t() + b.double()
    scale = float((A.double().abs() @ W.double().abs().t()).max())
    err = (C.double() - ref).abs()
    # fp16 result: half an ulp of the value (2^-11 relative) + fp32 accumulation noise
    assert float((err - ref.abs() * 2.0 ** -11).max()) <= 2e-6 * max(scale, 1.0), float(err.max())


def test_dynamic_row_count_leaves_tail_untouched(cuda_device):
    A = torch.randn(1000, 200, device=cuda_device)
    W = torch.randn(1536, 200, device=cuda_device) / 16
    b = torch.zeros(1536, device=cuda_device)
    C = _run("ttr_gemm_tf32_bias", A, W, b, m_valid=333)
    # whole 128-row tiles are stored by the copy engine: rows of the last touched tile (333..383) may be
    # written (they are never read by the callers), everything beyond stays untouched
    assert torch.isnan(C[384:]).all() and not torch.isnan(C[:333]).any()
    ref = A[:333].double() @ W.double().t()
    assert float((C[:333].double() - ref).abs().max()) < 2e-2


def test_tf32_operand_rounding_probe(cuda_device):
    """Records how kind::tf32 treats fp32 operands (printed, only sanity-pinned): C[i, :] =
    hw(1 + i * 2^-14) * hw(1) for a single non-zero k, under both tensor-map data types."""
    A = torch.zeros(128, 32, device=cuda_device)
    A[:, 0] = 1.0 + torch.arange(128, device=cuda_device, dtype=torch.float32) * 2.0 ** -14
    W = torch.zeros(128, 32, device=cuda_device)
    W[:, 0] = 1.0
    b = torch.zeros(128, device=cuda_device)
    res = {}
    for flags, name in ((0, "TFLOAT32 map"), (2, "FLOAT32 map")):
        _lib.call_nostream("ttr_debug_set_flags", flags)
        try:
            C = _run("ttr_gemm_tf32_bias", A, W, b)
        finally:
            _lib.call_nostream("ttr_debug_set_flags", 0)
        steps = ((C[:, 0].double() - 1.0) * 2.0 ** 14).round().long().cpu().tolist()
        res[name] = steps
        print(f"\n[tf32 probe] {name}: hw(1 + i*2^-14) in units of 2^-14, i=0..47: {steps[:48]}")
        assert all(abs(v - i) <= 16 for i, v in enumerate(steps))          # within one tf32 ulp (2^-10)
    print("[tf32 probe] identical across map types:", res["TFLOAT32 map"] == res["FLOAT32 map"])


@pytest.mark.parametrize("M,N1,N2,lda_extra,ldb_extra", [(64, 128, 128, 0, 0), (1000, 1536, 200, 0, 0), (5000, 768, 256, 768, 256),
                                                          (333, 96, 12, 0, 0), (40000, 1536, 512, 0, 0), (31, 48, 20, 48, 0)])
def test_tn_weight_gradient_gemm_matches_fp64(cuda_device, M, N1, N2, lda_extra, ldb_extra):
    """C = A[:M]^T B[:M] (MN-major tcgen05 + split-K + TMA reduce-add) vs fp64, incl. column blocks of wider
    matrices, a dynamic row count and accumulate."""
    g = torch.Generator(device=cuda_device).manual_seed(M + N1 + N2)
    m_bound = M + 77
    A = torch.randn(m_bound, N1 + lda_extra, device=cuda_device, generator=g) * 0.1
    Bm = torch.randn(m_bound, N2 + ldb_extra, device=cuda_device, generator=g)
    A[M:] = float("nan")                      # rows beyond m_valid are garbage until the tail is zeroed
    Bm[M:] = float("nan")
    mv = torch.tensor([M], dtype=torch.int32, device=cuda_device)
    for t in (A, Bm):
        _lib.call("ttr_zero_tail_rows", t, m_bound, mv, t.shape[1])
    a_view, b_view = A[:, lda_extra:], Bm[:, ldb_extra:]     # column blocks with pitch > width
    C = torch.full((N1, N2), float("nan"), device=cuda_device)
    _lib.call("ttr_gemm_tn_tf32", a_view, A.shape[1], b_view, Bm.shape[1], C, N2, m_bound, mv, N1, N2, 0)
    ref = a_view[:M].double().t() @ b_view[:M].double()
    scale = float((a_view[:M].double().abs().t() @ b_view[:M].double().abs()).max())
    err = float((C.double() - ref).abs().max())
    assert err <= 1.5e-3 * scale, (err, scale)
    _lib.call("ttr_gemm_tn_tf32", a_view, A.shape[1], b_view, Bm.shape[1], C, N2, m_bound, mv, N1, N2, 1)
    assert float((C.double() - 2 * ref).abs().max()) <= 3e-3 * scale


@pytest.mark.parametrize("M,N,K,mv", [(1000, 1536, 512, None), (513, 1536, 512, 300), (256, 256, 320, None), (5000, 1536, 512, 4321),
                                      (300, 520, 264, None), (129, 104, 456, 1), (20000, 1536, 512, None),
                                      (20000, 1536, 200, None), (5000, 1536, 200, 4097), (300, 1536, 200, None), (700, 520, 256, None),
                                      (129, 104, 8, 1), (40000, 768, 200, 39000)])
def test_f16_pair_gemm_matches_fp64_and_the_single_cta_kernel(cuda_device, M, N, K, mv):
    """Every K runs CTA pairs (tcgen05.mma.cta_group::2, 256 x 256 tiles, each CTA loads half of A's and half of W's rows);
    K <= 256 keeps each CTA's W rows resident and streams only A (debug bit 19: stream W too); debug bit 30 selects the
    single-CTA 128 x 128 kernels.  Same operands, same k order -> same fp16 results; rows at or beyond the device-side
    row count stay untouched."""
    g = torch.Generator(device=cuda_device).manual_seed(M + N + K)
    A = (torch.randn(M, K, device=cuda_device, generator=g) * 0.4).half()
    W = ((torch.rand(N, K, device=cuda_device, generator=g) - 0.5) / 8).half()
    b = torch.randn(N, device=cuda_device, generator=g) * 0.05
    mvt = None if mv is None else torch.tensor([mv], dtype=torch.int32, device=cuda_device)
    out = {}
    for name, flags in (("pair", 0), ("single", 1 << 30), ("pair_streamed_w", 1 << 19), ("pair_coalesced_store", 1 << 18),
                        ("pair_a_multicast", 1 << 16)):
        C = torch.full((M, N), 7.0, dtype=torch.float16, device=cuda_device)
        _lib.call_nostream("ttr_debug_set_flags", flags)
        try:
            _lib.call("ttr_gemm_f16_bias", A, W, b, C, M, mvt, N, K)
            torch.cuda.synchronize()
        finally:
            _lib.call_nostream("ttr_debug_set_flags", 0)
        out[name] = C
    rows = M if mv is None else mv
    ref = A[:rows].double() @ W.double().t() + b.double()
    scale = float((A.double().abs() @ W.double().abs().t()).max())
    err = (out["pair"][:rows].double() - ref).abs()
    assert float((err - ref.abs() * 2.0 ** -11).max()) <= 2e-6 * max(scale, 1.0), float(err.max())
    assert torch.equal(out["pair"][:rows], out["single"][:rows])
    assert torch.equal(out["pair"][:rows], out["pair_streamed_w"][:rows])
    assert torch.equal(out["pair"][:rows], out["pair_a_multicast"][:rows])      # debug bit 16: clusters of four, A multicast to two pairs
    assert torch.equal(out["pair"][:rows], out["pair_coalesced_store"][:rows])  # debug bit 18: staged coalesced stores instead of TMA stores
    tail_lo = (rows + 255) // 256 * 256                    # rows of partially valid tiles may be written (caller-owned, never read)
    if tail_lo < M:
        assert (out["pair"][tail_lo:] == 7.0).all()
