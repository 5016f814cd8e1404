#!/usr/bin/env python3
"""Generate tests/golden/frontend_search.npz by calling the UNMODIFIED `/search` handler of
`/root/reference/frontend/main.py` (`:102-210`) in this container.

    python oracle/make_golden_frontend.py     # needs /root/reference (absent on the GPU box)

TEST INFRASTRUCTURE — not part of the product path.

`frontend/main.py` imports `chromadb` (absent here, a third-party store the reference does not
vendor; its requirements pin none).  Everything else it needs is installed (fastapi, pydantic,
sklearn).  A stand-in module named `chromadb` is put into `sys.modules` for the import.  It offers
the three calls the file makes — `PersistentClient(path)`, `get_or_create_collection(name)`,
`collection.count()` / `collection.query(query_embeddings, n_results)` — over the documents and
embeddings of the artefact directory, and answers `query` by EXHAUSTIVE float32 squared-L2
distance, ascending (Chroma's default `hnsw:space` is `l2`, whose hnswlib kernel returns
sum((a-b)^2) in float32; an HNSW index at ef >= n returns exactly this list).  So the fixture
pins every line of the handler itself — both branches, `1 - dist`, the re-transform of the
candidate strings, sklearn's cosine, the alpha blend, the stable sort, the response body — and
leaves ONE thing taken from documentation: that the store's distance is squared L2.

The artefact directory is produced by the reference writer (`backend/main.py:92-153`) from seeded
synthetic triplets, exactly as in `make_golden.py::case_artifacts_eval` (H = 256 so that the
tcgen05 recurrence is the kernel under test on the GPU side).
"""
from __future__ import annotations

import importlib.util
import json
import os
import pickle
import sys
import tempfile
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
REF = Path(os.environ.get("TTR_REFERENCE", "/root/reference"))
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(REF / "backend"))

from twotowermlretrieval_b200 import synth  # noqa: E402

OUT = ROOT / "tests" / "golden"


def chroma_stand_in(documents, embeddings):
    """`chromadb` with the surface `frontend/main.py:29,72-74,153-156` touches."""
    emb = np.ascontiguousarray(embeddings, dtype=np.float32)

    class Collection:
        def count(self):
            return len(documents)

        def query(self, query_embeddings, n_results):
            docs, dists = [], []
            for q in query_embeddings:
                q = np.asarray(q, dtype=np.float32)
                diff = emb - q[None, :]
                d = np.einsum("nd,nd->n", diff, diff, dtype=np.float32)          # hnswlib L2Sqr, float32
                order = np.argsort(d, kind="stable")[:n_results]
                docs.append([documents[i] for i in order])
                dists.append([float(d[i]) for i in order])
            return {"documents": docs, "distances": dists}

    class PersistentClient:
        def __init__(self, path=None):
            self.path = path

        def get_or_create_collection(self, name):
            return Collection()

    mod = types.ModuleType("chromadb")
    mod.PersistentClient = PersistentClient
    return mod


def main():
    torch.manual_seed(0)
    torch.set_num_threads(1)
    sys.modules.setdefault("fastparquet", types.ModuleType("fastparquet"))    # backend/main.py -> data_loader import only
    os.environ.setdefault("WANDB_MODE", "disabled")
    import main as ref_main                       # /root/reference/backend/main.py (the writer), unmodified
    import model as refmodel
    import tokenizer as ref_tok
    words = ["the", "machine", "learning", "deep", "neural", "network", "data", "text", "image", "video",
             "language", "natural", "vision", "computer", "model", "layer", "search", "query", "document",
             "retrieval", "vector", "index", "tower", "embedding", "train", "loss", "cosine", "score",
             "rank", "passage", "and", "of", "in", ".", ",", "?"]
    w2i = {w: i for i, w in enumerate(words)}
    cfg = {"HIDDEN_DIM": 256, "RNN_TYPE": "GRU", "NUM_LAYERS": 2, "BIDIRECTIONAL": True, "DROPOUT": 0.0,
           "MARGIN": 0.5, "NORMALIZE_OUTPUT": True, "EMBED_DIM": 12, "BATCH_SIZE": 64}
    rng = np.random.default_rng(303)
    pool = [" ".join(rng.choice(words[1:33], size=int(rng.integers(5, 16)))) + " ." for _ in range(140)]
    queries = [" ".join(rng.choice(words[1:30], size=int(rng.integers(2, 6)))) for _ in range(40)]
    triplets = [(queries[i % 40], pool[i], pool[int(rng.integers(0, 140))]) for i in range(140)]
    probes = ["machine learning model", "deep neural network layer and vision", "zzz qqq", "the of in",
              "vector search index ?", "loss", "image video text data language natural", "the the", ""]
    alphas = [0.5, 1.0, 0.2, 0.0, 0.85]
    with tempfile.TemporaryDirectory() as td:
        td = Path(td)
        art = td / "artifacts" / "run-f"
        (td / "frontend").mkdir()
        (td / "frontend" / "config.json").write_text(json.dumps({"ARTIFACTS_PATH": str(art)}))
        w2i_path = td / "word_to_idx.pkl"
        with open(w2i_path, "wb") as f:
            pickle.dump(w2i, f)
        tok = ref_tok.PretrainedTokenizer(str(w2i_path))
        full = dict(cfg, VOCAB_SIZE=tok.vocab_size())
        sd = synth.make_state_dict(full, seed=41, table_seed=42)
        model = refmodel.TwoTowerModel(full, sd["query_encoder.embedding.weight"])
        model.load_state_dict({k: torch.tensor(v) for k, v in sd.items()})
        model.device = torch.device("cpu")
        ref_main.save_inference_artifacts(art, model, dict(cfg, WORD_TO_IDX_PATH=str(w2i_path)), tok,
                                          {"train": triplets, "validation": triplets[:20]})
        with open(art / "documents.pkl", "rb") as f:
            documents = pickle.load(f)
        doc_emb = np.load(art / "document_embeddings.npy")
        with open(art / "tfidf_artifacts.pkl", "rb") as f:
            tf = pickle.load(f)
        mat = tf["matrix"].tocsr()
        mat.sort_indices()
        sys.modules["chromadb"] = chroma_stand_in(documents, doc_emb)
        cwd = os.getcwd()
        os.chdir(td)                               # the file opens 'frontend/config.json' relative to the CWD (:28)
        try:
            spec = importlib.util.spec_from_file_location("ref_frontend_main", REF / "frontend" / "main.py")
            fe = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(fe)            # unmodified; builds QueryInferencer on the directory above
            fe.inferencer.device = torch.device("cpu")
            fe.inferencer.model.to("cpu")
            responses, q_embs = [], []
            for q in probes:
                try:
                    q_embs.append(fe.inferencer.get_query_embedding(q))
                except RuntimeError:               # "the the": ids [0, 0] -> length 0 -> pack_padded_sequence raises
                    q_embs.append(np.full(cfg["HIDDEN_DIM"], np.nan, np.float32))
                for a in alphas:
                    try:
                        r = fe.search(fe.QueryInput(query=q, alpha=a))
                        responses.append({"query": q, "alpha": a, "raises": False, "response": r})
                    except RuntimeError as e:
                        responses.append({"query": q, "alpha": a, "raises": True, "response": None})
        finally:
            os.chdir(cwd)
            del sys.modules["chromadb"]
    blob = {"cfg": json.dumps(full), "words": json.dumps(words), "triplets": json.dumps(triplets),
            "documents": json.dumps(documents), "doc_emb": doc_emb, "weight_seeds": np.array([41, 42]),
            "tfidf_indptr": mat.indptr, "tfidf_indices": mat.indices, "tfidf_data": mat.data,
            "tfidf_vocab": json.dumps({k: int(v) for k, v in tf["vectorizer"].vocabulary_.items()}),
            "tfidf_idf": tf["vectorizer"].idf_, "tfidf_shape": np.array(mat.shape),
            "probes": json.dumps(probes), "alphas": np.array(alphas), "probe_emb": np.stack(q_embs),
            "responses": json.dumps(responses)}
    np.savez_compressed(OUT / "frontend_search.npz", **blob)
    n_raise = sum(r["raises"] for r in responses)
    print(f"frontend /search ok: {len(responses)} responses ({n_raise} raise), {len(documents)} documents")
    print(json.dumps(responses[0]["response"]["results"][:2])[:400])


if __name__ == "__main__":
    main()
