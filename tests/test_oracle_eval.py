"""CPU: pin the evaluator / artefact-format restatements to the fixture produced by running the
unmodified reference writer, evaluators and inferencer (oracle/make_golden.py::case_artifacts_eval)."""
import json
import pickle

import numpy as np
import torch

from conftest import load_golden
from oracle import torch_path, towers_numpy as onp
from twotowermlretrieval_b200 import synth
from twotowermlretrieval_b200.data import TripletDataset, collate_fn, triplet_batches
from twotowermlretrieval_b200.tokenizer import PretrainedTokenizer


def _setup(tmp_path):
    g = load_golden("artifacts_eval")
    words = json.loads(str(g["words"]))
    p = tmp_path / "word_to_idx.pkl"
    with open(p, "wb") as fh:
        pickle.dump({w: i for i, w in enumerate(words)}, fh)
    tok = PretrainedTokenizer(str(p))
    cfg = g["cfg"]
    sd = synth.make_state_dict(cfg, seed=int(g["weight_seeds"][0]), table_seed=int(g["weight_seeds"][1]))
    triplets = [tuple(t) for t in json.loads(str(g["triplets"]))]
    return g, tok, cfg, sd, triplets


def test_reference_writer_fixture_is_reproduced_by_the_oracle(tmp_path):
    g, tok, cfg, sd, triplets = _setup(tmp_path)
    docs = json.loads(str(g["documents"]))
    assert set(docs) == {d for _, p, n in triplets for d in (p, n)}
    assert json.loads(str(g["saved_cfg"]))["VOCAB_SIZE"] == tok.vocab_size() == cfg["VOCAB_SIZE"]
    assert json.loads(str(g["saved_cfg"]))["EMBED_DIM"] == cfg["EMBED_DIM"]
    assert sorted(json.loads(str(g["state_keys"]))) == sorted(sd.keys())
    # document embeddings: the bulk-encode restatement (batch 64 loop) on the same document order
    sdt = torch_path.to_torch_state(sd)
    with torch.no_grad():
        emb = torch_path.bulk_encode_documents(sdt, cfg, [tok.encode(d) for d in docs], batch_size=64).numpy()
    np.testing.assert_allclose(emb, g["doc_emb"], rtol=1e-5, atol=1e-6)
    # TF-IDF: sklearn with the writer's parameters reproduces the stored matrix
    from sklearn.feature_extraction.text import TfidfVectorizer
    mat = TfidfVectorizer(stop_words="english", max_features=20000).fit_transform(docs).tocsr()
    mat.sort_indices()
    np.testing.assert_array_equal(mat.indptr, g["tfidf_indptr"])
    np.testing.assert_array_equal(mat.indices, g["tfidf_indices"])
    np.testing.assert_array_equal(mat.data, g["tfidf_data"])
    # the reference reader's output on that directory
    probe = json.loads(str(g["probe_queries"]))
    with torch.no_grad():
        for q, want in zip(probe, g["probe_emb"]):
            ids = tok.encode(q)
            got = torch_path.encoder_forward(sdt, "query_encoder", torch.tensor([ids]), cfg).numpy()[0]
            np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-6)


def test_evaluator_restatements_match_reference_metrics(tmp_path):
    g, tok, cfg, sd, triplets = _setup(tmp_path)
    val = triplets[: int(g["n_val"])]
    sdt = torch_path.to_torch_state(sd)
    qs, ps, ns, loss = [], [], [], 0.0
    batches = list(triplet_batches(val, tok, batch_size=16))
    with torch.no_grad():
        for q, p, n in batches:
            qe = torch_path.encoder_forward(sdt, "query_encoder", q, cfg)
            pe = torch_path.encoder_forward(sdt, "doc_encoder", p, cfg)
            ne = torch_path.encoder_forward(sdt, "doc_encoder", n, cfg)
            loss += float(torch_path.triplet_loss_cosine(qe, pe, ne, margin=cfg["MARGIN"]))
            qs.append(qe); ps.append(pe)
    m, _ = onp.batch_eval_metrics(torch.cat(qs).numpy(), torch.cat(ps).numpy())
    want = json.loads(str(g["batch_metrics"]))
    assert m == want or all(abs(m[k] - want[k]) < 1e-12 for k in want)
    assert abs(loss / len(batches) - float(g["batch_loss"])) < 1e-6
    # corpus evaluator: unique queries against all documents
    q2p, all_docs = {}, set()
    for q, p, n in val:
        q2p.setdefault(q, set()).add(p)
        all_docs.update((p, n))
    queries, docs = list(q2p), sorted(all_docs)
    with torch.no_grad():
        de = torch_path.bulk_encode_documents(sdt, cfg, [tok.encode(d) for d in docs], batch_size=64).numpy()
        qe = np.stack([torch_path.encoder_forward(sdt, "query_encoder", torch.tensor([tok.encode(q)]), cfg).numpy()[0]
                       for q in queries])
    cm = onp.corpus_eval_metrics(qe, de, queries, q2p, docs)
    want = json.loads(str(g["corpus_metrics"]))
    assert all(abs(cm[k] - want[k]) < 1e-12 for k in want), (cm, want)


def test_collate_matches_reference_contract(tmp_path):
    g, tok, cfg, sd, triplets = _setup(tmp_path)
    ds = TripletDataset(triplets[:5], tok)
    q, p, n = collate_fn([ds[i] for i in range(5)])
    q2, p2, n2 = next(triplet_batches(triplets[:5], tok, batch_size=5))
    assert torch.equal(q, q2) and torch.equal(p, p2) and torch.equal(n, n2)
    assert q.dtype == torch.int64 and q.shape[1] == max(len(tok.encode(t[0])) for t in triplets[:5])
    assert (q[0, len(tok.encode(triplets[0][0])):] == 0).all()


def test_frontend_search_restatement_matches_the_unmodified_handler():
    """`oracle/make_golden_frontend.py` ran `/root/reference/frontend/main.py::search` (`:102-210`) itself, over a
    stand-in store answering by exhaustive float32 squared L2.  The restatement must give the same response body."""
    from sklearn.feature_extraction.text import TfidfVectorizer
    g = load_golden("frontend_search")
    docs = json.loads(str(g["documents"]))
    vocab = json.loads(str(g["tfidf_vocab"]))
    vec = TfidfVectorizer()                                   # the reference's default vectoriser (`main.py:139`)
    vec.vocabulary_ = vocab
    vec.idf_ = g["tfidf_idf"]
    probes = json.loads(str(g["probes"]))
    emb_of = dict(zip(probes, g["probe_emb"]))
    n_checked = 0
    for item in json.loads(str(g["responses"])):
        q, alpha = item["query"], item["alpha"]
        if item["raises"]:                                    # "the the": zero-length sequence (quirk #2)
            assert alpha != 0.0 and np.isnan(emb_of[q]).all()
            continue
        qrow = vec.transform([q]).tocsr()
        qrow.sort_indices()
        q_emb = np.zeros(g["doc_emb"].shape[1], np.float32) if np.isnan(emb_of[q]).all() else emb_of[q]
        got = onp.frontend_search(q_emb, g["doc_emb"], docs, g["tfidf_indptr"], g["tfidf_indices"], g["tfidf_data"],
                                  qrow.indices, qrow.data, alpha)
        want = item["response"]["results"]
        assert item["response"]["query"] == q and item["response"]["alpha"] == alpha
        assert len(got) == len(want), (q, alpha)
        for r, (a, b) in enumerate(zip(got, want)):
            assert b["rank"] == r + 1 and b["id"] == f"result-{r + 1}"
            assert abs(a["score"] - b["score"]) <= 2e-6, (q, alpha, r)
            if a["doc"] != b["doc"]:                          # float32-distance ties in the store may order differently
                # (the query "" is the zero vector: every document is at distance |d|^2 = 1 +- 3e-7, one 140-way tie)
                assert abs(a["dense_score"] - b["dense_score"]) <= 2e-6 and abs(a["tfidf_score"] - b["tfidf_score"]) <= 2e-6
                assert q == "" or any(abs(w["score"] - a["score"]) <= 2e-6 and w["doc"] == a["doc"] for w in want)
                continue
            assert abs(a["dense_score"] - b["dense_score"]) <= 2e-6      # the store's distance is float32
            assert abs(a["tfidf_score"] - b["tfidf_score"]) <= 1e-12
            n_checked += 1
    assert n_checked > 250
