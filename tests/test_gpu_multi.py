"""GPU, 2 / 4 / 8 ranks (needs that many devices; skipped otherwise): row-sharded search / hybrid rerank / keyword
branch over the peer-memory exchange (and the NCCL all-gather fallback) == single-device result, and data-parallel
training == single-device training on the concatenated batch: the ALL-REDUCED FLAT GRADIENT BUCKET is compared
tightly with the single-rank global-batch gradient (a wrong all-reduce scale cannot hide there), the parameters
after two fused clip + Adam steps only as a secondary check.  Plus: launches follow the tensors' device, not the
thread's current device."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import scipy.sparse as sp
        from twotowermlretrieval_b200 import TwoTowerModel, synth, triplet_loss_cosine
        from twotowermlretrieval_b200.index import CsrF64, ShardedIndex, hybrid_rerank, search_topk, shard_bounds
        from twotowermlretrieval_b200.optim import FusedClipAdam
        from twotowermlretrieval_b200.search_service import SearchService
        out = {}
        groups = [dist.new_group([r]) for r in range(world)]      # new_group must be called by all ranks for every group
        # ---- sharded search + hybrid
        N, F = 30011, 300
        D = torch.tensor(synth.make_unit_rows(N, 256, seed=3), device=dev)
        indptr, indices, data = synth.make_tfidf_csr(N, n_features=F, mean_nnz=8, seed=2)
        full_csr = CsrF64.from_arrays(indptr, indices, data, dev)
        lo, hi = shard_bounds(N, world, rank)
        idx = ShardedIndex(D[lo:hi].contiguous(), lo, N, tfidf_local=full_csr.row_slice(lo, hi))
        idx_nccl = ShardedIndex(D[lo:hi].contiguous(), lo, N, tfidf_local=full_csr.row_slice(lo, hi), peer_memory=False)
        qptr = np.arange(0, 4 * 5 + 1, 4)
        rng = np.random.default_rng(1)
        qidx = np.concatenate([np.sort(rng.choice(F, 4, replace=False)) for _ in range(5)])
        qcsr = CsrF64.from_arrays(qptr, qidx, np.full(20, 0.5), dev)
        for B in (3, 40, 200):
            Q = torch.tensor(synth.make_unit_rows(B, 256, seed=40 + B), device=dev)
            s, i = idx.search(Q, 50)
            s1, i1 = search_topk(Q, D, 50)
            out[f"search{B}"] = bool(torch.equal(i, i1) and torch.allclose(s, s1, atol=1e-6))
            s2, i2 = idx_nccl.search(Q, 50)
            out[f"search{B}"] &= bool(torch.equal(i2, i1) and torch.equal(s2, s))
            for _ in range(5):                      # buffer alternation / flag steps over repeated calls
                s3, i3 = idx.search(Q, 50)
            out[f"search{B}"] &= bool(torch.equal(i3, i1) and torch.equal(s3, s))
        # a slow rank: the others run ahead into the next step's local scan and must wait at the flags
        Q = torch.tensor(synth.make_unit_rows(40, 256, seed=77), device=dev)
        s1, i1 = search_topk(Q, D, 50)
        ok = True
        for it in range(4):
            if rank == it % world:
                torch.cuda._sleep(int(2e8))         # ~0.1 s of GPU time on one rank
            s, i = idx.search(Q, 50)
            ok &= bool(torch.equal(i, i1))
        out["search_skewed"] = ok
        # deferred merges (side stream, three rotating buffers): a stream of 12 different batches enqueued without
        # waiting for any result, one rank delayed every few steps, interleaved with synchronous calls and a hybrid
        # call on the same exchange -> every result equals the single-device one
        Qs = [torch.tensor(synth.make_unit_rows(33, 256, seed=300 + j), device=dev) for j in range(12)]
        want = [search_topk(q_, D, 50) for q_ in Qs]
        for rnd in range(2):
            pend = []
            for j, q_ in enumerate(Qs):
                if rank == j % world and j % 3 == 0:
                    torch.cuda._sleep(int(5e7))
                if j == 7 and rnd == 1:
                    s_sync, i_sync = idx.search(Qs[0], 50)              # a synchronous call in the middle of the stream
                    ok &= bool(torch.equal(i_sync, want[0][1]))
                pend.append(idx.search_deferred(q_, 50))
            for (ws_, wi_), p_ in zip(want, pend):
                s_, i_ = p_.result()
                ok &= bool(torch.equal(i_, wi_) and torch.allclose(s_, ws_, atol=1e-6))
        out["search_deferred"] = ok
        Q = torch.tensor(synth.make_unit_rows(5, 256, seed=9), device=dev)
        h = idx.search_hybrid(Q, qcsr, alpha=0.4, k=50, top_n=10)
        s1, i1 = search_topk(Q, D, 50)
        h1 = hybrid_rerank(i1, s1, 0.4, docs_csr=full_csr, q_csr=qcsr, top_n=10)
        out["hybrid"] = bool(torch.equal(h["idx"], h1["idx"]) and torch.equal(h["final"], h1["final"]))
        h2 = idx_nccl.search_hybrid(Q, qcsr, alpha=0.4, k=50, top_n=10)
        out["hybrid"] &= bool(torch.equal(h2["idx"], h1["idx"]) and torch.equal(h2["final"], h1["final"]))
        out["peer_memory"] = bool(idx.peer_memory)
        # ---- alpha == 0 keyword branch of /search on the sharded index == the single-shard answer
        docs_txt = [f"doc {j}" for j in range(N)]
        q_row = sp.csr_matrix((np.full(4, 0.5), qidx[:4], np.array([0, 4])), shape=(1, F))

        def service(index):
            svc = SearchService.__new__(SearchService)
            svc.index, svc.documents, svc.device = index, docs_txt, dev
            return svc

        one = ShardedIndex(D, 0, N, group=groups[rank], tfidf_local=full_csr)
        kw_sharded = service(idx)._keyword(q_row)
        kw_single = service(one)._keyword(q_row)
        out["keyword"] = bool(kw_sharded == kw_single and len(kw_single) > 0)
        # ---- data-parallel training: world ranks x 8 triplets == 1 rank x 8*world triplets
        cfg = synth.default_config(vocab_size=3000, embed_dim=200)
        cfg["DROPOUT"] = 0.0
        sd = synth.make_state_dict(cfg, seed=0, table_seed=1)
        G = 8 * world
        q_ids, _ = synth.make_tokens(G, "query", 3000, seed=11)
        p_ids, _ = synth.make_tokens(G, "passage", 3000, seed=12, lengths=np.random.default_rng(12).integers(8, 30, G))
        n_ids, _ = synth.make_tokens(G, "passage", 3000, seed=13, lengths=np.random.default_rng(13).integers(8, 30, G))

        def fresh():
            m = TwoTowerModel(cfg, sd["query_encoder.embedding.weight"])
            m.load_state_dict({k: torch.tensor(v) for k, v in sd.items()})
            return m.to(dev).train()

        def backward(m, sl):
            loss = triplet_loss_cosine((m.encode_query(torch.tensor(q_ids[sl], device=dev)),
                                        m.encode_document(torch.tensor(p_ids[sl], device=dev)),
                                        m.encode_document(torch.tensor(n_ids[sl], device=dev))), margin=0.5)
            loss.backward()
            return loss

        m_dp, m_one = fresh(), fresh()
        opt_dp = FusedClipAdam(m_dp, lr=1e-3, max_norm=1.0)
        opt_one = FusedClipAdam(m_one, lr=1e-3, max_norm=1.0, process_group=groups[rank])
        sl = slice(rank * 8, rank * 8 + 8)
        # (1) the all-reduced gradient bucket vs the global-batch gradient, per tensor
        opt_dp.zero_grad(); opt_one.zero_grad()
        backward(m_dp, sl); backward(m_one, slice(0, G))
        g_dp = m_dp.flat_grads().clone()
        dist.all_reduce(g_dp, op=dist.ReduceOp.SUM)
        g_dp /= world
        g_one = m_one.flat_grads()
        worst, worst_name = 0.0, ""
        for name, (o, n) in m_one._slices.items():
            a, b = g_dp[o:o + n], g_one[o:o + n]
            rel = float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))
            if rel > worst:
                worst, worst_name = rel, name
        out["dp_grad_rel_err"] = worst
        out["dp_grad_worst_tensor"] = worst_name
        out["dp_grad_norms"] = (float(g_dp.norm()), float(g_one.norm()))
        # (2) secondary: two fused steps
        for _ in range(2):
            for m, opt, s_ in ((m_dp, opt_dp, sl), (m_one, opt_one, slice(0, G))):
                opt.zero_grad()
                backward(m, s_)
                opt.step()
        a, b = m_dp.flat_params(), m_one.flat_params()
        out["dp_max_param_diff"] = float((a - b).abs().max())
        out["dp_frac_off"] = float(((a - b).abs() > 1e-4).float().mean())
        out["dp_norms"] = (float(opt_dp.last_grad_norm), float(opt_one.last_grad_norm))
        torch.cuda.synchronize()
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_search_hybrid_keyword_and_dp_training(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 1000) + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    res = dict(q.get(timeout=900) for _ in range(world))
    [p.join(120) for p in procs]
    print(f"world {world}: rank 0 -> {res[0]}")
    for rank in range(world):
        r = res[rank]
        assert r["search3"] and r["search40"] and r["search200"] and r["search_skewed"] and r["hybrid"] and r["keyword"], r
        assert r["search_deferred"], r
        assert r["peer_memory"], "symmetric-memory exchange was not active"
        # all-reduce(SUM) / world of the per-rank mean-loss gradients == gradient of the global-batch mean loss, up to
        # the fp32/tf32 summation order of the kernels (different row tilings): 2e-3 of each tensor's largest entry.
        # A wrong reduction scale would show up as an O(1) relative error here.
        assert r["dp_grad_rel_err"] <= 2e-3, r
        assert abs(r["dp_grad_norms"][0] - r["dp_grad_norms"][1]) <= 1e-3 * r["dp_grad_norms"][1], r
        # secondary: clip + Adam on both sides.  Adam's first steps move every weight by lr * sign(g): an element whose
        # gradient is below the summation noise can take the other sign, so single elements may differ by up to
        # 2 steps * 2 lr while all but a vanishing fraction agree.
        assert r["dp_max_param_diff"] <= 4.1e-3, r
        assert r["dp_frac_off"] < 1e-3, r
        assert abs(r["dp_norms"][0] - r["dp_norms"][1]) < 1e-3 * r["dp_norms"][1], r


def test_launches_follow_the_tensor_device_not_the_current_device():
    """`device=` arguments must work when they differ from the thread's current device."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from twotowermlretrieval_b200 import TwoTowerModel, synth, _lib
    from twotowermlretrieval_b200.index import search_topk
    torch.cuda.set_device(0)
    d0, d1 = torch.device("cuda:0"), torch.device("cuda:1")
    D = torch.tensor(synth.make_unit_rows(20000, 256, seed=3))
    Q = torch.tensor(synth.make_unit_rows(40, 256, seed=4))
    s0, i0 = search_topk(Q.to(d0), D.to(d0), 50)
    s1, i1 = search_topk(Q.to(d1), D.to(d1), 50)               # current device is still 0
    assert s1.device == d1 and torch.equal(i0.cpu(), i1.cpu()) and torch.equal(s0.cpu(), s1.cpu())
    cfg = synth.default_config(vocab_size=3000, embed_dim=200)
    sd = synth.make_state_dict(cfg, seed=0, table_seed=1)
    ids, _ = synth.make_tokens(64, "passage", 3000, seed=12)
    embs = []
    for dev in (d0, d1):
        m = TwoTowerModel(cfg, sd["query_encoder.embedding.weight"])
        m.load_state_dict({k: torch.tensor(v) for k, v in sd.items()})
        m.to(dev).eval()
        with torch.no_grad():
            embs.append(m.encode_document(torch.tensor(ids, device=dev)).cpu())
    assert torch.cuda.current_device() == 0
    assert torch.allclose(embs[0], embs[1], atol=1e-6)
    with pytest.raises(_lib.TTRError):
        search_topk(Q.to(d0), D.to(d1), 50)                    # operands on different devices are rejected
