"""CPU: pin both oracle restatements to the fixtures generated from the unmodified
reference (oracle/make_golden.py), and to each other."""
import json

import numpy as np
import pytest
import torch

from conftest import golden_weights, load_golden
from oracle import torch_path, towers_numpy as onp
from twotowermlretrieval_b200 import synth

SMALL = ["small_bi2", "small_uni1", "small_bi1_trainable_table", "small_uni2"]


@pytest.mark.parametrize("name", SMALL)
def test_numpy_oracle_matches_reference_fixture(name):
    g = load_golden(name)
    sd, cfg = golden_weights(g), g["cfg"]
    for key, tower in (("q", "query_encoder"), ("p", "doc_encoder"), ("n", "doc_encoder")):
        out = onp.encoder_forward(sd, tower, g[key], cfg, dtype=np.float64)
        np.testing.assert_allclose(out, g[f"{key}_emb"], rtol=2e-5, atol=2e-6)
    loss = onp.triplet_loss_cosine(g["q_emb"], g["p_emb"], g["n_emb"], margin=cfg["MARGIN"])
    assert abs(loss - float(g["loss"])) < 1e-6


@pytest.mark.parametrize("name", SMALL)
def test_torch_oracle_matches_reference_fixture(name):
    g = load_golden(name)
    cfg = g["cfg"]
    sd = torch_path.to_torch_state(golden_weights(g))
    with torch.no_grad():
        for key, tower in (("q", "query_encoder"), ("p", "doc_encoder"), ("n", "doc_encoder")):
            out = torch_path.encoder_forward(sd, tower, torch.tensor(g[key]), cfg)
            np.testing.assert_allclose(out.numpy(), g[f"{key}_emb"], rtol=1e-5, atol=1e-6)
        # per-layer decomposition with all-ones masks == fused nn.GRU
        if cfg["NUM_LAYERS"] > 1:
            dirs = 2 if cfg["BIDIRECTIONAL"] else 1
            x = torch.tensor(g["p"])
            ones = [torch.ones(x.shape[0], x.shape[1], dirs * cfg["HIDDEN_DIM"])]
            out = torch_path.encoder_forward(sd, "doc_encoder", x, cfg, dropout_masks=ones)
            np.testing.assert_allclose(out.numpy(), g["p_emb"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("name", ["small_bi2", "small_uni1", "small_uni2"])
def test_torch_oracle_train_step_matches_reference(name):
    g = load_golden(name)
    cfg = g["cfg"]
    sd = torch_path.to_torch_state(golden_weights(g), requires_grad=True)
    loss, gnorm = torch_path.train_step(sd, {}, cfg, torch.tensor(g["q"]), torch.tensor(g["p"]), torch.tensor(g["n"]))
    assert abs(loss - float(g["train_loss"])) < 1e-6
    assert abs(gnorm - float(g["grad_norm"])) < 1e-5
    for k, t in sd.items():
        if f"a::{k}" in g:
            np.testing.assert_allclose(t.detach().numpy(), g[f"a::{k}"], rtol=1e-5, atol=1e-6)


def test_cfgdims_fixture_against_both_oracles():
    g = load_golden("cfgdims")
    cfg = g["cfg"]
    sd_np = synth.make_state_dict(cfg, seed=0, table_seed=1)
    sd = torch_path.to_torch_state(sd_np)
    with torch.no_grad():
        pe = torch_path.encoder_forward(sd, "doc_encoder", torch.tensor(g["p"]), cfg).numpy()
    np.testing.assert_allclose(pe, g["p_emb"], rtol=1e-5, atol=1e-6)
    # numpy float64 oracle on a few rows (explicit loops are slow at H=256)
    out = onp.encoder_forward(sd_np, "query_encoder", g["q"][:3], cfg, dtype=np.float64)
    np.testing.assert_allclose(out, g["q_emb"][:3], rtol=1e-4, atol=2e-6)
    # two live training steps
    sdg = torch_path.to_torch_state(sd_np, requires_grad=True)
    st = {}
    for i in range(2):
        loss, gn = torch_path.train_step(sdg, st, cfg, torch.tensor(g["q"]), torch.tensor(g["p"]), torch.tensor(g["n"]))
        assert abs(loss - g["train_losses"][i]) < 1e-5
        assert abs(gn - g["grad_norms"][i]) < 1e-4
    for k, t in sdg.items():
        if f"ah::{k}" in g:
            np.testing.assert_allclose(t.detach().reshape(-1)[:64].numpy(), g[f"ah::{k}"], rtol=1e-4, atol=1e-6)


def test_zero_length_row_raises_like_reference():
    g = load_golden("small_bi2")
    x = g["q"].copy()
    x[3, :] = 0
    with pytest.raises(RuntimeError):
        onp.encoder_forward(golden_weights(g), "query_encoder", x, g["cfg"])
    with pytest.raises(RuntimeError):
        torch_path.encoder_forward(torch_path.to_torch_state(golden_weights(g)), "query_encoder",
                                   torch.tensor(x), g["cfg"])


def test_quirk1_first_nnz_positions():
    """[5,0,7,9] encodes exactly like [5,0,7]."""
    g = load_golden("small_bi2")
    sd, cfg = golden_weights(g), g["cfg"]
    a = onp.encoder_forward(sd, "query_encoder", np.array([[5, 0, 7, 9]]), cfg)
    b = onp.encoder_forward(sd, "query_encoder", np.array([[5, 0, 7, 0]]), cfg)
    c = onp.encoder_forward(sd, "query_encoder", np.array([[5, 0, 7, 9, 0, 0]]), cfg)
    np.testing.assert_allclose(a, c, atol=1e-12)
    assert np.abs(a - b).max() > 1e-4            # [5,0,7,0] has length 2 -> differs


def test_search_fixture():
    g = load_golden("search")
    D = synth.make_unit_rows(int(g["n_docs"]), int(g["dim"]), seed=int(g["doc_seed"]))
    Q = synth.make_unit_rows(5, int(g["dim"]), seed=int(g["query_seed"]))
    s, i = onp.cosine_topk(Q, D, 50, dtype=np.float64)
    np.testing.assert_allclose(s, g["scores"], rtol=1e-5, atol=1e-6)
    # identical indices except where neighbouring scores tie inside fp32 rounding
    diff = i != g["idx"]
    if diff.any():
        gap = np.abs(np.diff(g["scores"], axis=1)).min()
        assert gap < 1e-6, "index mismatch without a near-tie"
    vt, it = torch_path.cosine_topk(torch.tensor(Q), torch.tensor(D), 50)
    assert (it.numpy() == g["idx"]).all()


def test_hybrid_and_inferencer_fixture():
    g = load_golden("inferencer_hybrid")
    cfg = g["cfg"]
    words = json.loads(str(g["words"]))
    docs = json.loads(str(g["docs"]))
    queries = json.loads(str(g["queries"]))
    w2i = {w: i for i, w in enumerate(words)}
    unk = len(w2i)
    import re
    enc = lambda s: [w2i.get(t, unk) for t in re.findall(r"\w+|[.,!?;]", s.lower())]
    sd = golden_weights(g)
    from sklearn.feature_extraction.text import TfidfVectorizer
    vec = TfidfVectorizer(stop_words="english", max_features=10000)
    mat = vec.fit_transform(docs)
    demb = np.stack([onp.encoder_forward(sd, "query_encoder", np.array([enc(d)]), cfg)[0] for d in docs])
    np.testing.assert_allclose(demb, g["doc_emb"], rtol=1e-4, atol=2e-6)
    for qi, q in enumerate(queries):
        ids = enc(q)
        if g["raises"][qi]:
            with pytest.raises(RuntimeError):
                onp.encoder_forward(sd, "query_encoder", np.array([ids]), cfg)
            continue
        qe = onp.encoder_forward(sd, "query_encoder", np.array([ids]), cfg)[0]
        np.testing.assert_allclose(qe, g["query_emb"][qi], rtol=1e-4, atol=2e-6)
        dense = (demb @ qe) / (np.linalg.norm(demb, axis=1) * np.linalg.norm(qe))
        tf = (mat @ vec.transform([q]).T).toarray()[:, 0]
        top, sc = onp.hybrid_search_simple(dense, tf, float(g["alpha"]), 10)
        np.testing.assert_allclose(sc, g["res_score"][qi], rtol=1e-5, atol=1e-6)
        same = top == g["res_idx"][qi]
        assert same.all() or np.abs(np.diff(g["res_score"][qi])).min() < 1e-6


def test_frontend_rerank_restatement_properties():
    indptr, indices, data = synth.make_tfidf_csr(200, n_features=50, mean_nnz=6, seed=9)
    rng = np.random.default_rng(3)
    cand = rng.choice(200, size=50, replace=False)
    cos = np.sort(rng.random(50))[::-1]
    qi, qv = np.array([3, 7, 20], np.int32), np.array([0.5, 0.5, np.sqrt(0.5)])
    # alpha=1 keeps dense order; semantic = 2cos-1 under Chroma's default l2 space
    order, final, sem, tf = onp.hybrid_rerank_frontend(cand, cos, indptr, indices, data, qi, qv, 1.0)
    assert (order == np.arange(10)).all()
    np.testing.assert_allclose(sem, 2 * cos[:10] - 1)
    # empty TF-IDF query row -> zero keyword scores (frontend/main.py:173-175)
    order, final, sem, tf = onp.hybrid_rerank_frontend(cand, cos, indptr, indices, data, [], [], 0.3)
    assert (tf == 0).all() and (order == np.arange(10)).all()
    # tfidf against scipy
    import scipy.sparse as sp
    M = sp.csr_matrix((data, indices, indptr), shape=(200, 50))
    qd = np.zeros(50); qd[qi] = qv
    order, final, sem, tf = onp.hybrid_rerank_frontend(cand, cos, indptr, indices, data, qi, qv, 0.4, top_n=50)
    np.testing.assert_allclose(tf, (M[cand] @ qd)[order], atol=1e-12)
    assert (np.diff(final) <= 1e-15).all()


def test_adam_and_clip_match_torch():
    rng = np.random.default_rng(0)
    p0 = rng.standard_normal(50).astype(np.float32)
    g0 = rng.standard_normal(50).astype(np.float32) * 3
    t = torch.tensor(p0, requires_grad=True)
    opt = torch.optim.Adam([t], lr=5e-5)
    p, m, v = p0.astype(np.float64), np.zeros(50), np.zeros(50)
    for step in (1, 2, 3):
        t.grad = torch.tensor(g0)
        tot = torch.nn.utils.clip_grad_norm_([t], 1.0)
        opt.step()
        (gc,), tn = onp.clip_grad_norm([g0], 1.0)
        assert abs(tn - float(tot)) < 1e-4
        p, m, v = onp.adam_step(p, gc.astype(np.float64), m, v, step, 5e-5)
        np.testing.assert_allclose(p, t.detach().numpy(), rtol=1e-5, atol=1e-7)


def test_config1_fixture_first_steps_match_the_torch_restatement():
    """tests/golden/config1_epoch.npz was produced by the unmodified reference loop (oracle/make_golden_config1.py);
    the functional restatement `oracle.torch_path.train_step` reproduces its first two steps on the same batches."""
    import torch
    from oracle import torch_path
    from oracle.make_golden_config1 import config1_inputs
    from twotowermlretrieval_b200.data import pad_rows
    g = load_golden("config1_epoch")
    assert int(g["n_steps"]) == 157 and len(g["losses"]) == 157 and np.isfinite(g["losses"]).all()
    cfg, words, triplets, perm, sd = config1_inputs()
    w2i = {w: i for i, w in enumerate(words)}
    enc = lambda s: [w2i[t] for t in s.split()]
    sdt = torch_path.to_torch_state(sd, requires_grad=True)
    st = {}
    for i in range(2):
        rows = [triplets[int(j)] for j in perm[i * 64:(i + 1) * 64]]
        q, p, n = (pad_rows([enc(r[c]) for r in rows]) for c in range(3))
        loss, gn = torch_path.train_step(sdt, st, cfg, q, p, n)
        assert abs(loss - g["losses"][i]) < 2e-5, (i, loss, g["losses"][i])
        assert abs(gn - g["grad_norms"][i]) < 1e-3 * g["grad_norms"][i]
