"""GPU parity: fused cosine top-k / merge kernels vs the oracle (evaluators.py:185-186)."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from gpu_util import check_topk
from oracle import towers_numpy as onp
from twotowermlretrieval_b200 import synth
from twotowermlretrieval_b200.index import search_topk, topk_merge

pytestmark = pytest.mark.gpu


def test_golden_fixture(cuda_device):
    g = load_golden("search")
    D = synth.make_unit_rows(int(g["n_docs"]), 256, seed=int(g["doc_seed"]))
    Q = synth.make_unit_rows(5, 256, seed=int(g["query_seed"]))
    s, i = search_topk(torch.tensor(Q, device=cuda_device), torch.tensor(D, device=cuda_device), 50)
    np.testing.assert_allclose(s.cpu().numpy(), g["scores"], rtol=1e-3, atol=1e-4)
    assert (i.cpu().numpy() == g["idx"]).mean() > 0.95      # B=5 runs the tf32 path; check_topk verifies every swap is a near-tie
    check_topk(s, i, Q, D, 50)


@pytest.mark.parametrize("B", [1, 2, 3, 4, 7, 8, 9, 21])
@pytest.mark.parametrize("N", [1, 31, 32, 33, 257, 5000])
def test_small_and_ragged_shapes(cuda_device, B, N):
    D = synth.make_unit_rows(N, 256, seed=100 + N)
    Q = synth.make_unit_rows(B, 256, seed=200 + B)
    k = 50
    s, i = search_topk(torch.tensor(Q, device=cuda_device), torch.tensor(D, device=cuda_device), k, row_offset=1000)
    kk = min(k, N)
    check_topk(s[:, :kk], i[:, :kk], Q, D, kk, row_offset=1000)
    if N < k:                                # unfilled slots are flagged, not garbage
        assert (i[:, N:] == -1).all() and torch.isinf(s[:, N:]).all()


@pytest.mark.parametrize("k", [1, 5, 10, 50, 64])
def test_k_values(cuda_device, k):
    D = synth.make_unit_rows(3000, 256, seed=7)
    Q = synth.make_unit_rows(6, 256, seed=8)
    s, i = search_topk(torch.tensor(Q, device=cuda_device), torch.tensor(D, device=cuda_device), k)
    check_topk(s, i, Q, D, k)          # B=6 runs the tcgen05 path: near-tie swaps are verified, not forbidden
    assert (np.diff(s.cpu().numpy(), axis=1) <= 0).all()


def test_adversarial_orders_and_ties(cuda_device):
    rng = np.random.default_rng(0)
    q = synth.make_unit_rows(1, 256, seed=1)
    # ascending scores: every document beats the running threshold -> constant compaction
    N = 6000
    D = rng.standard_normal((N, 256)).astype(np.float32) * 0.01
    D += np.linspace(-1, 1, N, dtype=np.float32)[:, None] * q
    s, i = search_topk(torch.tensor(q, device=cuda_device), torch.tensor(D, device=cuda_device), 50)
    check_topk(s, i, q, D, 50)
    # all-equal documents: exact ties resolve to the lowest indices, in order
    D2 = np.repeat(synth.make_unit_rows(1, 256, seed=2), 4000, axis=0)
    s, i = search_topk(torch.tensor(q, device=cuda_device), torch.tensor(D2, device=cuda_device), 50)
    assert i.cpu().tolist()[0] == list(range(50))
    # duplicated blocks: each winner appears with all its copies, lowest index first
    base = synth.make_unit_rows(500, 256, seed=3)
    D3 = np.concatenate([base, base, base])
    s, i = search_topk(torch.tensor(q, device=cuda_device), torch.tensor(D3, device=cuda_device), 30)
    ii = i.cpu().numpy()[0]
    assert (ii[0::3] + 500 == ii[1::3]).all() and (ii[1::3] + 500 == ii[2::3]).all()


@pytest.mark.parametrize("B,N", [(9, 31), (16, 1000), (33, 5000), (128, 20000), (130, 4097), (256, 30000), (300, 10000)])
def test_tcgen05_path_vs_oracle(cuda_device, B, N):
    """Batches > 4 run the tcgen05 kernel (tf32 operands rounded to nearest by the TMA engine)."""
    D = synth.make_unit_rows(N, 256, seed=300 + N)
    Q = synth.make_unit_rows(B, 256, seed=400 + B)
    k = 50
    s, i = search_topk(torch.tensor(Q, device=cuda_device), torch.tensor(D, device=cuda_device), k, row_offset=7)
    kk = min(k, N)
    swaps = check_topk(s[:, :kk], i[:, :kk], Q, D, kk, row_offset=7)   # every swap is verified to be a near-tie
    assert swaps <= 0.10 * B * kk, f"{swaps} rank swaps"
    assert (np.diff(s.cpu().numpy()[:, :kk], axis=1) <= 0).all()


def test_tcgen05_and_streaming_paths_agree(cuda_device):
    from twotowermlretrieval_b200 import _lib
    D = torch.tensor(synth.make_unit_rows(50000, 256, seed=41), device=cuda_device)
    Q = torch.tensor(synth.make_unit_rows(40, 256, seed=42), device=cuda_device)
    s_mma, i_mma = search_topk(Q, D, 50)
    _lib.call_nostream("ttr_debug_set_flags", 4)
    try:
        s_st, i_st = search_topk(Q, D, 50)
    finally:
        _lib.call_nostream("ttr_debug_set_flags", 0)
    assert torch.allclose(s_mma, s_st, rtol=1e-3, atol=1e-4)
    assert (i_mma == i_st).float().mean() > 0.9


def test_fused_sample_and_two_pass_paths_agree(cuda_device):
    """B <= 128 runs sample phase, bound selection (grid barrier) and main pass in one cooperative launch;
    debug bit 21 forces the separate sample / merge / seed launches.  Both are exact, so they agree bit for bit
    (the bound only prunes), also with fewer than k sample candidates per query and with exact ties."""
    from twotowermlretrieval_b200 import _lib
    for N, B in ((200_000, 128), (150_000, 5), (400_000, 77)):       # < 4 M documents: fused by default
        D = torch.tensor(synth.make_unit_rows(N, 256, seed=50 + B), device=cuda_device)
        D[90_000:90_200] = D[300:500]
        Q = torch.tensor(synth.make_unit_rows(B, 256, seed=60 + B), device=cuda_device)
        s_a, i_a = search_topk(Q, D, 50)
        _lib.call_nostream("ttr_debug_set_flags", 1 << 21)
        try:
            s_b, i_b = search_topk(Q, D, 50)
        finally:
            _lib.call_nostream("ttr_debug_set_flags", 0)
        assert torch.equal(s_a, s_b) and torch.equal(i_a, i_b)
        for _ in range(3):                                        # repeatable (the barrier counters are re-armed per call)
            s_c, i_c = search_topk(Q, D, 50)
            assert torch.equal(s_a, s_c) and torch.equal(i_a, i_c)
    check_topk(s_a[:6], i_a[:6], Q[:6].cpu().numpy(), D.cpu().numpy(), 50)


def test_select_merge_unstaged_path(cuda_device):
    """The select-merge stages candidate keys in shared memory; sets too large for that are re-read
    from L2 each radix round (debug flag bit 9 forces that path).  Same exact result either way."""
    from twotowermlretrieval_b200 import _lib
    D = torch.tensor(synth.make_unit_rows(60000, 256, seed=43), device=cuda_device)
    D[20000:20300] = D[100:400]                                   # exact score ties across document slices
    Q = torch.tensor(synth.make_unit_rows(130, 256, seed=44), device=cuda_device)
    s_a, i_a = search_topk(Q, D, 50)
    _lib.call_nostream("ttr_debug_set_flags", 512)
    try:
        s_b, i_b = search_topk(Q, D, 50)
    finally:
        _lib.call_nostream("ttr_debug_set_flags", 0)
    assert torch.equal(s_a, s_b) and torch.equal(i_a, i_b)
    check_topk(s_a[:8], i_a[:8], Q[:8].cpu().numpy(), D.cpu().numpy(), 50)


def test_tcgen05_adversarial(cuda_device):
    rng = np.random.default_rng(1)
    Q = synth.make_unit_rows(16, 256, seed=5)
    N = 9000
    D = rng.standard_normal((N, 256)).astype(np.float32) * 0.01
    D += np.linspace(-1, 1, N, dtype=np.float32)[:, None] * Q[0]       # ascending for query 0
    s, i = search_topk(torch.tensor(Q, device=cuda_device), torch.tensor(D, device=cuda_device), 50)
    check_topk(s, i, Q, D, 50)
    D2 = np.repeat(synth.make_unit_rows(1, 256, seed=2), 5000, axis=0)  # exact ties -> lowest ids
    s, i = search_topk(torch.tensor(Q, device=cuda_device), torch.tensor(D2, device=cuda_device), 50)
    assert (i.cpu().numpy() == np.arange(50)[None, :]).all()


def test_merge_kernel_matches_oracle(cuda_device):
    rng = np.random.default_rng(5)
    P, B, kin, k = 7, 9, 50, 50
    s = np.sort(rng.standard_normal((P, B, kin)).astype(np.float32), axis=2)[:, :, ::-1].copy()
    i = rng.permutation(P * B * kin).reshape(P, B, kin).astype(np.int64)
    s[2, :, 10:] = s[3, :, 10:]                                  # cross-list ties
    ms, mi = topk_merge(torch.tensor(s, device=cuda_device), torch.tensor(i, device=cuda_device), k)
    cs = s.transpose(1, 0, 2).reshape(B, -1)
    ci = i.transpose(1, 0, 2).reshape(B, -1)
    order = np.lexsort((ci, -cs), axis=1)[:, :k]
    np.testing.assert_array_equal(mi.cpu().numpy(), np.take_along_axis(ci, order, 1))
    np.testing.assert_array_equal(ms.cpu().numpy(), np.take_along_axis(cs, order, 1))


def test_sharded_equals_unsharded(cuda_device):
    """Row shards + merge == one matrix (the multi-GPU scheme, emulated on one device)."""
    D = synth.make_unit_rows(10007, 256, seed=11)
    Q = synth.make_unit_rows(5, 256, seed=12)
    Dd, Qd = torch.tensor(D, device=cuda_device), torch.tensor(Q, device=cuda_device)
    s_full, i_full = search_topk(Qd, Dd, 50)
    from twotowermlretrieval_b200.index import shard_bounds
    parts = []
    for r in range(3):
        lo, hi = shard_bounds(D.shape[0], 3, r)
        parts.append(search_topk(Qd, Dd[lo:hi].contiguous(), 50, row_offset=lo))
    ms, mi = topk_merge(torch.stack([p[0] for p in parts]), torch.stack([p[1] for p in parts]), 50)
    assert torch.equal(mi, i_full) and torch.equal(ms, s_full)
    # one process: the deferred entry point is the plain search behind the same handle type
    from twotowermlretrieval_b200.index import PendingSearch, ShardedIndex
    pend = ShardedIndex(Dd).search_deferred(Qd, 50)
    assert isinstance(pend, PendingSearch)
    s_d, i_d = pend.result()
    assert torch.equal(i_d, i_full) and torch.equal(s_d, s_full)
    assert pend.result()[0] is s_d                                 # idempotent


def test_million_docs_properties(cuda_device):
    """BASELINE configs[1] size: N=1M, B=1 and 8 — checked against torch on the same device in
    chunks (size-independent properties: sorted, unique, every reported score is the true dot,
    nothing outside the list beats the k-th)."""
    gen = torch.Generator(device=cuda_device).manual_seed(3)
    D = torch.nn.functional.normalize(torch.randn(1_000_000, 256, device=cuda_device, generator=gen), dim=1)
    Q = torch.nn.functional.normalize(torch.randn(8, 256, device=cuda_device, generator=gen), dim=1)
    Q = torch.nn.functional.normalize(torch.randn(256, 256, device=cuda_device, generator=gen), dim=1)
    for B in (1, 8, 256):
        s, i = search_topk(Q[:B], D, 50)
        full = (Q[:B].double() @ D.double().t())
        ref_s, ref_i = torch.topk(full, 50, dim=1)
        assert torch.allclose(s.double(), ref_s, rtol=1e-3, atol=1e-4)
        assert (torch.diff(s, dim=1) <= 0).all()
        got = torch.gather(full, 1, i)
        assert torch.allclose(got, ref_s, rtol=1e-3, atol=1e-4)        # every returned doc scores like the true rank
        assert (i == ref_i).float().mean() > (0.99 if B <= 4 else 0.9)


def _fp64_topk_on_device(Q, D, k, chunk=1 << 19):
    """Chunked fp64 matmul + topk on the device (the reference's evaluators.py:185-186 in fp64), ties by lower index;
    returns (scores, idx) of the exact top-k."""
    Qd = Q.double()
    bs = torch.empty(Q.shape[0], 0, dtype=torch.float64, device=Q.device)
    bi = torch.empty(Q.shape[0], 0, dtype=torch.int64, device=Q.device)
    for lo in range(0, D.shape[0], chunk):
        sc = Qd @ D[lo:lo + chunk].double().t()
        s, i = torch.topk(sc, min(k, sc.shape[1]), dim=1)
        bs, bi = torch.cat([bs, s], 1), torch.cat([bi, i + lo], 1)
        if bs.shape[1] > 8 * k:
            o = torch.argsort(bs, dim=1, descending=True, stable=True)[:, :k]
            bs, bi = torch.gather(bs, 1, o), torch.gather(bi, 1, o)
    o = torch.argsort(bi, dim=1, stable=True)
    bs, bi = torch.gather(bs, 1, o), torch.gather(bi, 1, o)
    o = torch.argsort(bs, dim=1, descending=True, stable=True)[:, :k]
    return torch.gather(bs, 1, o), torch.gather(bi, 1, o)


def _check_vs_fp64(s, i, Q, D, k, rtol=1e-3, atol=1e-4):
    """north_star tolerance: scores within 1e-3 relative (+1e-4 absolute), indices identical except at ties in it."""
    ref_s, ref_i = _fp64_topk_on_device(Q, D, k)
    tol = rtol * ref_s.abs() + atol
    assert ((s.double() - ref_s).abs() <= tol).all(), float((s.double() - ref_s).abs().max())
    assert (torch.diff(s, dim=1) <= 0).all()
    true = (D[i.reshape(-1)].double().view(Q.shape[0], k, -1) * Q.double().unsqueeze(1)).sum(-1)
    mism = i != ref_i
    assert ((true - ref_s).abs() <= tol)[mism].all()          # a different document only where it ties the reference's
    srt = torch.sort(i, dim=1).values
    assert (srt[:, 1:] != srt[:, :-1]).all(), "duplicate indices"
    return float(mism.float().mean())


@pytest.mark.parametrize("B,n_check", [(128, 128), (256, 256), (4096, 256)])
def test_headline_corpus_8p8M_docs(cuda_device, B, n_check):
    """The bench shapes themselves: N = 8,841,823 (MS MARCO passage count) at B = 128 (two-pass path, 1,867 tiles per
    CTA), B = 256 (two query tiles, 3,734 tiles per CTA = two id segments) and B = 4096 (32 query tiles, 4 slices,
    34 segments), checked against chunked fp64 on the device (every 16th query at B = 4096)."""
    N = 8_841_823
    gen = torch.Generator(device=cuda_device).manual_seed(3)
    D = torch.empty(N, 256, device=cuda_device)
    for lo in range(0, N, 1 << 20):
        hi = min(N, lo + (1 << 20))
        D[lo:hi] = torch.nn.functional.normalize(torch.randn(hi - lo, 256, device=cuda_device, generator=gen), dim=1)
    Q = torch.nn.functional.normalize(torch.randn(B, 256, device=cuda_device, generator=gen), dim=1)
    s, i = search_topk(Q, D, 50)
    sel = torch.arange(0, B, B // n_check, device=cuda_device)
    frac = _check_vs_fp64(s[sel], i[sel], Q[sel], D, 50)
    assert frac < 0.1, frac
    s2, i2 = search_topk(Q, D, 50)                              # repeatable
    assert torch.equal(i, i2) and torch.equal(s, s2)


@pytest.mark.parametrize("B,N", [(128, 300_000), (256, 300_000), (300, 150_000), (1024, 100_000), (40, 70_001), (129, 3333),
                                 (512, 1_000_003)])
def test_id_segments_and_fast_reject(cuda_device, B, N):
    """B > 128 runs CTA PAIRS (tcgen05.mma.cta_group::2, 64-document tiles, each CTA loads half of every tile).
    A CTA numbers its candidates with 16 bits inside a segment and publishes its list at every segment boundary;
    debug bit 24 shrinks the segments to 16 tiles so a small corpus crosses dozens of boundaries (the production
    length is 2,048 tiles: B >= 256 on the full corpus).  Exact, so: identical to the default segmentation, to the
    two-pass path, and to the epilogue without the max-tree fast reject (bit 25)."""
    from twotowermlretrieval_b200 import _lib
    D = torch.tensor(synth.make_unit_rows(N, 256, seed=70 + B), device=cuda_device)
    D[N // 2:N // 2 + 300] = D[100:400].clone()               # exact ties across slices and segments
    Q = torch.tensor(synth.make_unit_rows(B, 256, seed=80 + B), device=cuda_device)
    s_a, i_a = search_topk(Q, D, 50)
    # bit 27: one CTA per query tile (round-1 layout) instead of CTA pairs for B > 128
    for flags in (1 << 24, (1 << 24) | (1 << 21), 1 << 25, (1 << 24) | (1 << 25), 1 << 21, 1 << 27, (1 << 27) | (1 << 24),
                  1 << 29, (1 << 29) | (1 << 24), (1 << 29) | (1 << 21),     # bit 29: no screening warps
                  -(1 << 31), -(1 << 31) | (1 << 24), -(1 << 31) | (1 << 21)):  # bit 31: no survivor-histogram bound
        _lib.call_nostream("ttr_debug_set_flags", flags)
        try:
            s_b, i_b = search_topk(Q, D, 50)
        finally:
            _lib.call_nostream("ttr_debug_set_flags", 0)
        assert torch.equal(s_a, s_b) and torch.equal(i_a, i_b), f"flags {flags:#x}"
    _check_vs_fp64(s_a[:16], i_a[:16], Q[:16], D, 50)


def test_adversarial_order_across_segments(cuda_device):
    """Ascending scores (every document beats the running bound -> constant compaction) with 16-tile segments."""
    from twotowermlretrieval_b200 import _lib
    rng = np.random.default_rng(2)
    Q = synth.make_unit_rows(130, 256, seed=6)
    N = 40000
    D = rng.standard_normal((N, 256)).astype(np.float32) * 0.01
    D += np.linspace(-1, 1, N, dtype=np.float32)[:, None] * Q[0]
    Qd, Dd = torch.tensor(Q, device=cuda_device), torch.tensor(D, device=cuda_device)
    _lib.call_nostream("ttr_debug_set_flags", 1 << 24)
    try:
        s, i = search_topk(Qd, Dd, 50)
    finally:
        _lib.call_nostream("ttr_debug_set_flags", 0)
    _check_vs_fp64(s[:8], i[:8], Qd[:8], Dd, 50)


@pytest.mark.parametrize("B", [128, 256, 40])
def test_survivor_histogram_bound_is_only_a_pruning_device(cuda_device, B):
    """The main pass counts every surviving candidate in a per-query histogram above the seeded bound; two idle warps per
    CTA turn it into a k-th-best lower bound shared by all CTAs (debug bit 31 switches it off).  It must never change
    the result — also when the score distribution defeats the bin placement: a heavy cluster far above the sample's
    maximum (everything lands in the open top bin), thousands of exact duplicates of the best documents (plateaus at
    the bound), scores that all fall into bin 0, and negative / tiny scores."""
    from twotowermlretrieval_b200 import _lib
    N = 600_000
    rng = np.random.default_rng(100 + B)
    Q = synth.make_unit_rows(B, 256, seed=7 + B)
    D = synth.make_unit_rows(N, 256, seed=8 + B)
    D[400_000:400_300] = 0.9 * Q[0] + 0.1 * D[400_000:400_300]           # 300 documents far above anything sampled (query 0)
    D[500_000:503_000] = D[123]                                          # 3,000 exact duplicates
    D[550_000:550_040] = Q[1 % B]                                        # 40 perfect matches of query 1: fewer than k
    D[10_000:20_000] *= 1e-3                                             # tiny scores
    D[30_000:40_000] *= -1.0
    Qd, Dd = torch.tensor(Q, device=cuda_device), torch.tensor(D, device=cuda_device)
    s_a, i_a = search_topk(Qd, Dd, 50)
    for flags in (-(1 << 31), 1 << 21, -(1 << 31) | (1 << 21)):
        _lib.call_nostream("ttr_debug_set_flags", flags)
        try:
            s_b, i_b = search_topk(Qd, Dd, 50)
        finally:
            _lib.call_nostream("ttr_debug_set_flags", 0)
        assert torch.equal(s_a, s_b) and torch.equal(i_a, i_b), f"flags {flags:#x}"
    _check_vs_fp64(s_a[:12], i_a[:12], Qd[:12], Dd, 50)
    for _ in range(3):
        s_c, i_c = search_topk(Qd, Dd, 50)
        assert torch.equal(s_a, s_c) and torch.equal(i_a, i_c)


def test_concurrent_streams_do_not_share_scratch(cuda_device):
    """re-entrant per CUDA stream — two streams search at the same time with their own workspaces."""
    D = torch.tensor(synth.make_unit_rows(400_000, 256, seed=91), device=cuda_device)
    Qs = [torch.tensor(synth.make_unit_rows(100, 256, seed=92 + j), device=cuda_device) for j in range(2)]
    want = [search_topk(q, D, 50) for q in Qs]
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(device=cuda_device) for _ in range(2)]
    got = [None, None]
    for rep in range(5):
        for j, st in enumerate(streams):
            with torch.cuda.stream(st):
                got[j] = search_topk(Qs[j], D, 50)
        torch.cuda.synchronize()
        for j in range(2):
            assert torch.equal(got[j][1], want[j][1]) and torch.equal(got[j][0], want[j][0])


def test_k_outside_kernel_range_is_a_clear_error(cuda_device):
    D = torch.tensor(synth.make_unit_rows(1000, 256, seed=1), device=cuda_device)
    Q = torch.tensor(synth.make_unit_rows(2, 256, seed=2), device=cuda_device)
    with pytest.raises(ValueError, match="outside"):
        search_topk(Q, D, 65)
