"""GPU parity: hybrid rerank kernel (frontend/main.py:158-198), QueryInferencer and
SimpleHybridRetriever against the fixture produced by running the reference classes."""
import json
import pickle

import numpy as np
import pytest
import torch

from conftest import golden_weights, load_golden
from oracle import towers_numpy as onp
from twotowermlretrieval_b200 import synth
from twotowermlretrieval_b200.index import CsrF64, ShardedIndex, hybrid_rerank, search_topk, tfidf_candidates

pytestmark = pytest.mark.gpu


def _queries_csr(rng, B, F, device):
    idx, val, ptr = [], [], [0]
    for b in range(B):
        k = 0 if b == 1 else int(rng.integers(2, 7))        # query 1 has no TF-IDF vocabulary hit
        c = np.sort(rng.choice(F, size=k, replace=False))
        v = rng.random(k) + 0.1
        v = v / np.linalg.norm(v) if k else v
        idx.append(c); val.append(v); ptr.append(ptr[-1] + k)
    return (CsrF64.from_arrays(np.array(ptr), np.concatenate(idx) if idx else [], np.concatenate(val), device),
            idx, val)


@pytest.mark.parametrize("space", ["l2", "cosine"])
@pytest.mark.parametrize("alpha", [1.0, 0.7, 0.3, 0.05])
def test_rerank_matches_frontend_restatement(cuda_device, alpha, space):
    rng = np.random.default_rng(17)
    N, F, B, kc = 3000, 400, 6, 50
    indptr, indices, data = synth.make_tfidf_csr(N, n_features=F, mean_nnz=12, seed=9)
    docs = CsrF64.from_arrays(indptr, indices, data, cuda_device)
    qcsr, qidx, qval = _queries_csr(rng, B, F, cuda_device)
    D = synth.make_unit_rows(N, 256, seed=3)
    Q = synth.make_unit_rows(B, 256, seed=4)
    s, i = search_topk(torch.tensor(Q, device=cuda_device), torch.tensor(D, device=cuda_device), kc)
    out = hybrid_rerank(i, s, alpha, docs_csr=docs, q_csr=qcsr, space=space, top_n=10)
    sc, ic = s.cpu().numpy(), i.cpu().numpy()
    for b in range(B):
        order, fin, sem, tf = onp.hybrid_rerank_frontend(ic[b], sc[b], indptr, indices, data, qidx[b], qval[b],
                                                        alpha, top_n=10, space=space)
        np.testing.assert_array_equal(out["pos"][b].cpu().numpy(), order)
        np.testing.assert_allclose(out["final"][b].cpu().numpy(), fin, rtol=0, atol=1e-15)
        np.testing.assert_allclose(out["tfidf"][b].cpu().numpy(), tf, rtol=0, atol=1e-15)
        np.testing.assert_array_equal(out["idx"][b].cpu().numpy(), ic[b][order])
    if alpha == 1.0:
        assert (out["pos"].cpu().numpy() == np.arange(10)).all()      # stable sort keeps dense order


def test_sharded_hybrid_single_rank_and_candidate_tfidf(cuda_device):
    rng = np.random.default_rng(3)
    N, F, B = 2000, 300, 4
    indptr, indices, data = synth.make_tfidf_csr(N, n_features=F, mean_nnz=10, seed=2)
    docs = CsrF64.from_arrays(indptr, indices, data, cuda_device)
    qcsr, qidx, qval = _queries_csr(rng, B, F, cuda_device)
    D = torch.tensor(synth.make_unit_rows(N, 256, seed=5), device=cuda_device)
    Q = torch.tensor(synth.make_unit_rows(B, 256, seed=6), device=cuda_device)
    idx = ShardedIndex(D, 0, N, tfidf_local=docs)
    out = idx.search_hybrid(Q, qcsr, alpha=0.4, k=50, top_n=10)
    s, i = search_topk(Q, D, 50)
    direct = hybrid_rerank(i, s, 0.4, docs_csr=docs, q_csr=qcsr, top_n=10)
    assert torch.equal(out["idx"], direct["idx"]) and torch.equal(out["final"], direct["final"])
    # a shard only scores the rows it owns
    half = docs.row_slice(1000, 2000)
    tf_half = tfidf_candidates(i, half, qcsr)
    tf_full = tfidf_candidates(i, docs, qcsr)
    own = (i >= 1000)
    assert torch.equal(tf_half[own], tf_full[own]) and (tf_half[~own] == 0).all()


def test_inferencer_and_simple_hybrid_match_reference_run(cuda_device, tmp_path):
    g = load_golden("inferencer_hybrid")
    cfg = g["cfg"]
    words = json.loads(str(g["words"]))
    docs = json.loads(str(g["docs"]))
    queries = json.loads(str(g["queries"]))
    art = tmp_path / "run"
    art.mkdir()
    with open(art / "word_to_idx.pkl", "wb") as f:
        pickle.dump({w: i for i, w in enumerate(words)}, f)
    torch.save({k: torch.tensor(v) for k, v in golden_weights(g).items()}, art / "model.pth")
    small = {k: v for k, v in cfg.items() if k != "VOCAB_SIZE"}
    (art / "config.json").write_text(json.dumps(small))
    from twotowermlretrieval_b200.query_inferencer import QueryInferencer
    from twotowermlretrieval_b200.simple_hybrid import SimpleHybridRetriever
    inf = QueryInferencer(str(art), device=cuda_device)
    assert inf.config["VOCAB_SIZE"] == len(words) + 1
    z = inf.get_query_embedding("")
    assert z.shape == (cfg["HIDDEN_DIM"],) and z.dtype == np.float32 and not z.any()     # query_inferencer.py:65-69
    r = SimpleHybridRetriever(str(art), alpha=float(g["alpha"]), device=cuda_device)
    r.fit(docs)
    np.testing.assert_allclose(r.doc_embeddings, g["doc_emb"], rtol=0, atol=3e-4)
    for qi, q in enumerate(queries):
        if g["raises"][qi]:
            with pytest.raises(RuntimeError):
                inf.get_query_embedding(q)
            continue
        e = inf.get_query_embedding(q)
        np.testing.assert_allclose(e, g["query_emb"][qi], rtol=0, atol=3e-4)
        res = r.search(q, top_k=10)
        got_idx = [docs.index(d) for d, _ in res]
        got_sc = np.array([s for _, s in res])
        np.testing.assert_allclose(got_sc, g["res_score"][qi], rtol=0, atol=5e-4)
        want = g["res_idx"][qi].tolist()
        assert got_idx == want or np.abs(np.diff(g["res_score"][qi])).min() < 1e-3
    batch = inf.encode_queries([q for qi, q in enumerate(queries) if not g["raises"][qi]] + [""])
    assert batch.shape[0] == 5 and not batch[-1].any()


def test_corpus_wide_blend_kernel_vs_numpy(cuda_device):
    """ttr_blend_topk == alpha*cos + (1-alpha)*tfidf over all docs, np.argsort()[::-1][:k] (simple_hybrid.py:57-60)."""
    import scipy.sparse as sp
    from twotowermlretrieval_b200 import _lib
    rng = np.random.default_rng(11)
    N, F, D = 20011, 500, 256
    docs = (rng.standard_normal((N, D)) * 0.3).astype(np.float32)           # not unit norm on purpose
    docs[17] = 0.0                                                          # zero row -> cosine 0
    indptr, indices, data = synth.make_tfidf_csr(N, n_features=F, mean_nnz=10, seed=4)
    M = sp.csr_matrix((data, indices, indptr), shape=(N, F))
    csr = CsrF64.from_arrays(indptr, indices, data, cuda_device)
    d_dev = torch.tensor(docs, device=cuda_device)
    lib = _lib.load()
    for alpha, k, qn in ((0.6, 10, 4), (0.0, 25, 3), (1.0, 64, 0)):
        q = rng.standard_normal(D).astype(np.float32)
        qi = np.sort(rng.choice(F, size=qn, replace=False)).astype(np.int32)
        qv = rng.random(qn) + 0.1
        qv = qv / np.linalg.norm(qv) if qn else qv
        qd = np.zeros(F); qd[qi] = qv
        dn = np.linalg.norm(docs, axis=1)
        dense = np.where(dn > 0, (docs @ q) / np.maximum(dn * np.linalg.norm(q), 1e-30), 0.0).astype(np.float32)
        comb = alpha * dense + (1 - alpha) * (M @ qd)
        top, sc = onp.hybrid_search_simple(dense, M @ qd, alpha, k)
        ws = torch.empty(int(lib.ttr_blend_topk_workspace_bytes(k)), dtype=torch.uint8, device=cuda_device)
        out_s = torch.empty(k, dtype=torch.float64, device=cuda_device)
        out_i = torch.empty(k, dtype=torch.int64, device=cuda_device)
        full = torch.empty(N, dtype=torch.float64, device=cuda_device)
        _lib.call("ttr_blend_topk", torch.tensor(q, device=cuda_device), float(np.linalg.norm(q)), d_dev, N, D,
                  csr.indptr, csr.indices, csr.data, torch.tensor(qi, device=cuda_device),
                  torch.tensor(qv, device=cuda_device), qn, alpha, k, out_s, out_i, full, ws)
        np.testing.assert_allclose(full.cpu().numpy(), comb, rtol=0, atol=2e-6)
        np.testing.assert_allclose(out_s.cpu().numpy(), sc, rtol=0, atol=2e-6)
        got = out_i.cpu().numpy()
        # exact agreement with the kernel's own combined vector, argsort()[::-1] tie order included
        own = full.cpu().numpy()
        assert (got == np.argsort(own, kind="stable")[::-1][:k]).all()
        assert (got == top).mean() > 0.9


def test_l2_space_with_a_zero_vector_query_and_unnormalised_docs(cuda_device):
    """a token-less query is the zero vector (query_inferencer.py:65-69); Chroma's squared-L2 distance is then
    |d|^2 and `semantic = 1 - dist` (frontend/main.py:162) is 0 for unit documents, not 2*0 - 1.  With `q_sqnorm` /
    `d_sqnorm` the kernel evaluates the general 1 - (|q|^2 + |d|^2 - 2 q.d); the restatement does the same."""
    rng = np.random.default_rng(23)
    N, F, kc = 500, 200, 50
    indptr, indices, data = synth.make_tfidf_csr(N, n_features=F, mean_nnz=10, seed=4)
    docs = CsrF64.from_arrays(indptr, indices, data, cuda_device)
    qcsr, qidx, qval = _queries_csr(rng, 3, F, cuda_device)
    D = synth.make_unit_rows(N, 256, seed=5) * np.linspace(0.5, 1.5, N, dtype=np.float32)[:, None]   # |d| != 1
    Q = synth.make_unit_rows(3, 256, seed=6)
    Q[0] = 0.0                                                                                       # token-less query
    Dd, Qd = torch.tensor(D, device=cuda_device), torch.tensor(Q, device=cuda_device)
    s, i = search_topk(Qd, Dd, kc)
    q_sq = (Qd.double() ** 2).sum(1)
    d_sq = (Dd[i].double() ** 2).sum(-1)
    out = hybrid_rerank(i, s, 0.6, docs_csr=docs, q_csr=qcsr, space="l2", top_n=10, q_sqnorm=q_sq, d_sqnorm=d_sq)
    sc, ic = s.cpu().numpy(), i.cpu().numpy()
    for b in range(3):
        order, fin, sem, tf = onp.hybrid_rerank_frontend(ic[b], sc[b], indptr, indices, data, qidx[b], qval[b], 0.6, top_n=10,
                                                        space="l2", q_sqnorm=float(q_sq[b]), d_sqnorm=d_sq[b].cpu().numpy())
        np.testing.assert_array_equal(out["pos"][b].cpu().numpy(), order)
        np.testing.assert_allclose(out["final"][b].cpu().numpy(), fin, rtol=0, atol=1e-15)
        np.testing.assert_allclose(out["semantic"][b].cpu().numpy(), sem, rtol=0, atol=1e-15)
    # zero query against unit documents: dense_score == 0 exactly
    U = torch.tensor(synth.make_unit_rows(N, 256, seed=7), device=cuda_device)
    s, i = search_topk(Qd[:1], U, kc)
    out = hybrid_rerank(i, s, 1.0, docs_csr=docs, q_csr=CsrF64(qcsr.indptr[:2].contiguous(), qcsr.indices, qcsr.data, 1, 0),
                        space="l2", top_n=10, q_sqnorm=q_sq[:1])
    assert (out["semantic"] == 0).all() and (out["final"] == 0).all()
