"""GPU parity for the rows SURVEY.md §8(f) lists next to the hot path: artefact round trip
(backend/main.py:92-153), the `/search` handler (frontend/main.py:102-210) and the evaluators
(backend/evaluators.py:9-209), against the fixture produced by running the unmodified reference
writer / evaluators / inferencer and against the oracle restatements."""
import json
import pickle

import numpy as np
import pytest
import torch

from conftest import load_golden
from gpu_util import model_from_numpy
from oracle import towers_numpy as onp
from twotowermlretrieval_b200 import synth
from twotowermlretrieval_b200.artifacts import load_corpus_artifacts, save_inference_artifacts
from twotowermlretrieval_b200.data import TripletDataset, collate_fn
from twotowermlretrieval_b200.evaluators import BatchEvaluator, CorpusEvaluator, positive_ranks
from twotowermlretrieval_b200.search_service import SearchService
from twotowermlretrieval_b200.tokenizer import PretrainedTokenizer

pytestmark = pytest.mark.gpu
REL_TOL = 1e-3


@pytest.fixture(scope="module")
def world(tmp_path_factory, cuda_device):
    g = load_golden("artifacts_eval")
    tmp = tmp_path_factory.mktemp("art")
    words = json.loads(str(g["words"]))
    w2i = tmp / "word_to_idx.pkl"
    with open(w2i, "wb") as fh:
        pickle.dump({w: i for i, w in enumerate(words)}, fh)
    tok = PretrainedTokenizer(str(w2i))
    cfg = g["cfg"]
    sd = synth.make_state_dict(cfg, seed=int(g["weight_seeds"][0]), table_seed=int(g["weight_seeds"][1]))
    model = model_from_numpy(cfg, sd, cuda_device).eval()
    model.device = cuda_device
    triplets = [tuple(t) for t in json.loads(str(g["triplets"]))]
    val = triplets[: int(g["n_val"])]
    art = tmp / "artifacts" / "run-b200"
    run_cfg = {k: v for k, v in cfg.items() if k not in ("VOCAB_SIZE",)}
    run_cfg["WORD_TO_IDX_PATH"] = str(w2i)
    save_inference_artifacts(art, model, run_cfg, tok, {"train": triplets, "validation": val})
    return dict(g=g, tok=tok, cfg=cfg, model=model, triplets=triplets, val=val, art=art, dev=cuda_device)


def test_artifact_directory_matches_reference_writer(world):
    g, art = world["g"], world["art"]
    assert sorted(p.name for p in art.iterdir()) == ["config.json", "document_embeddings.npy", "documents.pkl",
                                                    "model.pth", "tfidf_artifacts.pkl", "word_to_idx.pkl"]
    state = torch.load(art / "model.pth", map_location="cpu")
    assert list(state.keys()) == json.loads(str(g["state_keys"]))          # names AND order of the reference file
    saved = json.loads((art / "config.json").read_text())
    ref_saved = json.loads(str(g["saved_cfg"]))
    assert {k: saved[k] for k in ref_saved if k != "WORD_TO_IDX_PATH"} == {k: v for k, v in ref_saved.items() if k != "WORD_TO_IDX_PATH"}
    docs, emb, vec, mat = load_corpus_artifacts(art)
    ref_docs = json.loads(str(g["documents"]))
    assert set(docs) == set(ref_docs) and len(docs) == len(ref_docs)       # list(set(...)) order is per-process
    emb = np.asarray(emb)
    assert emb.dtype == np.float32 and emb.flags["C_CONTIGUOUS"] and emb.shape == g["doc_emb"].shape
    row_of = {d: i for i, d in enumerate(docs)}
    perm = np.array([row_of[d] for d in ref_docs])
    err = np.linalg.norm(emb[perm] - g["doc_emb"], axis=1) / np.linalg.norm(g["doc_emb"], axis=1)
    assert err.max() <= REL_TOL, err.max()
    # TF-IDF rows, aligned to the reference document order: same pattern; values to the last bit or two (sklearn
    # sums a row's squares in vocabulary-insertion order, which follows the per-process document order)
    mat = mat[perm].tocsr()
    mat.sort_indices()
    np.testing.assert_array_equal(mat.indptr, g["tfidf_indptr"])
    np.testing.assert_array_equal(mat.indices, g["tfidf_indices"])
    np.testing.assert_allclose(mat.data, g["tfidf_data"], rtol=1e-15, atol=0)
    assert {k: int(v) for k, v in vec.vocabulary_.items()} == json.loads(str(g["tfidf_vocab"]))
    with pytest.raises(FileNotFoundError):
        load_corpus_artifacts(art / "missing")


def test_query_inferencer_reads_the_directory_like_the_reference(world):
    from twotowermlretrieval_b200.query_inferencer import QueryInferencer
    g = world["g"]
    inf = QueryInferencer(str(world["art"]), device=world["dev"])
    for q, want in zip(json.loads(str(g["probe_queries"])), g["probe_emb"]):
        got = inf.get_query_embedding(q)
        assert got.dtype == np.float32 and got.shape == want.shape
        assert np.linalg.norm(got - want) <= REL_TOL * max(np.linalg.norm(want), 1e-12)


@pytest.mark.parametrize("alpha", [0.5, 1.0, 0.2, 0.0])
def test_search_service_matches_frontend_restatement(world, alpha):
    svc = SearchService(str(world["art"]), device=world["dev"])
    docs, emb, vec, mat = load_corpus_artifacts(world["art"])
    mat.sort_indices()
    for query in ["machine learning model", "deep neural network layer and vision", "zzz qqq", "the of in"]:
        out = svc.search(query, alpha)
        assert out["query"] == query and out["alpha"] == alpha
        qrow = vec.transform([query]).tocsr()
        qrow.sort_indices()
        q_emb = svc.inferencer.get_query_embedding(query)
        want = onp.frontend_search(q_emb, np.asarray(emb), docs, mat.indptr, mat.indices, mat.data, qrow.indices,
                                   qrow.data, alpha)
        res = out["results"]
        assert len(res) == len(want)
        for r, (got, ref) in enumerate(zip(res, want)):
            assert got["rank"] == r + 1 and got["id"] == f"result-{r + 1}"
            assert set(got) == {"rank", "id", "doc", "score", "dense_score", "tfidf_score"}
            assert abs(got["score"] - ref["score"]) <= 1e-5                 # same embeddings on both sides
            if got["doc"] != ref["doc"]:                                    # only equal-score neighbours may swap
                assert any(abs(got["score"] - w["score"]) <= 1e-6 and w["doc"] == got["doc"] for w in want)
            else:
                assert abs(got["dense_score"] - ref["dense_score"]) <= 1e-5
                assert abs(got["tfidf_score"] - ref["tfidf_score"]) <= 1e-12
        if alpha == 0.0:
            assert all(r["dense_score"] == 0.0 and r["score"] > 1e-5 for r in res)
        if query == "zzz qqq" and alpha not in (0.0,):                      # no TF-IDF vocabulary hit (frontend/main.py:173-175)
            assert all(r["tfidf_score"] == 0.0 for r in res)


def test_batch_evaluator_matches_reference(world):
    g, cfg, tok, model, dev = world["g"], world["cfg"], world["tok"], world["model"], world["dev"]
    loader = torch.utils.data.DataLoader(TripletDataset(world["val"], tok), batch_size=16, shuffle=False,
                                         collate_fn=collate_fn)
    metrics, loss = BatchEvaluator(top_k=[1, 5, 10]).evaluate(model, loader, dev, cfg)
    want = json.loads(str(g["batch_metrics"]))
    n = int(g["n_val"])
    assert set(metrics) == set(want)
    for k in (1, 5, 10):                       # tf32 projection: at most one near-tie rank flip per threshold
        assert abs(metrics[f"Recall@{k}"] - want[f"Recall@{k}"]) <= 1.0 / n + 1e-12
    assert abs(metrics["MRR"] - want["MRR"]) <= 0.02
    assert abs(loss - float(g["batch_loss"])) <= 3e-4
    # the rank kernel itself: exact against the restatement on the SAME embeddings
    with torch.no_grad():
        qs, ps = [], []
        for q, p, _ in loader:
            qs.append(model.encode_query(q.to(dev))); ps.append(model.encode_document(p.to(dev)))
        qe, pe = torch.cat(qs), torch.cat(ps)
    ranks = positive_ranks(qe, pe, torch.arange(qe.shape[0], device=dev)).cpu().numpy()
    m2, ranks_ref = onp.batch_eval_metrics(qe.cpu().numpy(), pe.cpu().numpy())
    np.testing.assert_array_equal(ranks, ranks_ref)
    assert all(abs(metrics[k] - m2[k]) < 1e-12 for k in m2)
    # duplicates: exact score ties resolve to the lower index, out-of-range targets are flagged
    dup = torch.cat([pe[:4], pe[:4]])
    r = positive_ranks(qe[:4], dup, torch.tensor([4, 5, 6, 99], device=dev)).cpu().tolist()
    r0 = positive_ranks(qe[:4], dup, torch.tensor([0, 1, 2, 3], device=dev)).cpu().tolist()
    assert r[:3] == [x + 1 for x in r0[:3]] and r[3] == -1


def test_corpus_evaluator_matches_reference(world):
    g, tok, model, dev = world["g"], world["tok"], world["model"], world["dev"]
    got = CorpusEvaluator(top_k=[1, 5, 10]).evaluate(model, world["val"], tok, dev)
    want = json.loads(str(g["corpus_metrics"]))
    assert set(got) == set(want)
    for k, v in want.items():
        assert abs(got[k] - v) <= 0.06, (k, got[k], v)       # one near-tie swap moves a mean over ~27 queries by <= 1/27
    with pytest.raises(RuntimeError):
        CorpusEvaluator(top_k=[500]).evaluate(model, world["val"], tok, dev)


def test_search_service_matches_the_unmodified_frontend_handler(tmp_path, cuda_device):
    """`tests/golden/frontend_search.npz` holds the responses of `/root/reference/frontend/main.py::search`
    (`:102-210`, unmodified) over a stand-in store answering by exhaustive float32 squared L2
    (`oracle/make_golden_frontend.py`).  Here the whole chain is ours: artefact writer on the GPU towers -> resident
    index -> `SearchService.search`.  Scores follow the embedding tolerance (1e-3 relative on unit vectors: dense_score =
    1 - |q - d|^2 moves by <= 2e-3); documents must be the reference's except where two scores are that close."""
    g = load_golden("frontend_search")
    words = json.loads(str(g["words"]))
    w2i = tmp_path / "word_to_idx.pkl"
    with open(w2i, "wb") as fh:
        pickle.dump({w: i for i, w in enumerate(words)}, fh)
    tok = PretrainedTokenizer(str(w2i))
    cfg = g["cfg"]
    sd = synth.make_state_dict(cfg, seed=int(g["weight_seeds"][0]), table_seed=int(g["weight_seeds"][1]))
    model = model_from_numpy(cfg, sd, cuda_device).eval()
    model.device = cuda_device
    triplets = [tuple(t) for t in json.loads(str(g["triplets"]))]
    art = tmp_path / "artifacts" / "run-fe"
    run_cfg = {k: v for k, v in cfg.items() if k != "VOCAB_SIZE"}
    run_cfg["WORD_TO_IDX_PATH"] = str(w2i)
    save_inference_artifacts(art, model, run_cfg, tok, {"train": triplets, "validation": triplets[:20]})
    svc = SearchService(str(art), device=cuda_device)
    TOL = 2e-3
    n_same, n_total = 0, 0
    for item in json.loads(str(g["responses"])):
        q, alpha = item["query"], item["alpha"]
        if item["raises"]:
            with pytest.raises(RuntimeError):
                svc.search(q, alpha)
            continue
        out = svc.search(q, alpha)
        want = item["response"]
        assert out["query"] == want["query"] and out["alpha"] == want["alpha"]
        res, ref = out["results"], want["results"]
        assert len(res) == len(ref), (q, alpha)
        for r, (a, b) in enumerate(zip(res, ref)):
            assert set(a) == set(b) and a["rank"] == b["rank"] and a["id"] == b["id"]
            assert abs(a["score"] - b["score"]) <= TOL * max(abs(alpha), 1e-9) + 1e-12, (q, alpha, r)
            n_total += 1
            if a["doc"] == b["doc"]:
                n_same += 1
                assert abs(a["dense_score"] - b["dense_score"]) <= (TOL if alpha != 0.0 else 0.0)
                assert abs(a["tfidf_score"] - b["tfidf_score"]) <= 1e-12
            elif q != "":                                   # "" is one 140-way tie at dense_score = 0 +- 3e-7
                assert any(w["doc"] == a["doc"] and abs(w["score"] - a["score"]) <= 2 * TOL for w in ref) or r >= len(ref) - 2
        if q == "" and alpha != 0.0:
            assert all(abs(a["dense_score"]) <= 1e-5 and a["tfidf_score"] == 0.0 for a in res)
    assert n_same >= 0.9 * n_total - 40, (n_same, n_total)
