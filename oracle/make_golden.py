#!/usr/bin/env python3
"""Generate tests/golden/*.npz by running the UNMODIFIED reference in this container.

    python oracle/make_golden.py            # needs /root/reference (absent on the GPU box)

Imports `/root/reference/backend/model.py`, `query_inferencer.py`, `simple_hybrid.py`
as they are (sys.path only; nothing is copied), feeds them the seeded inputs from
`twotowermlretrieval_b200.synth`, and stores inputs + outputs.  The fixtures pin both
oracle restatements (tests/test_oracle.py) and, on the GPU, the CUDA path
(tests/test_gpu_*.py).  Re-running must reproduce the committed files bit-for-bit on the
same torch build (2.11.0+cu128 CPU).
"""
from __future__ import annotations

import json
import os
import pickle
import sys
import tempfile
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
REF = Path(os.environ.get("TTR_REFERENCE", "/root/reference"))
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(REF / "backend"))

from twotowermlretrieval_b200 import synth  # noqa: E402

OUT = ROOT / "tests" / "golden"


def ref_model(cfg, sd_np, pretrained: bool):
    import model as refmodel                      # /root/reference/backend/model.py
    table = sd_np["query_encoder.embedding.weight"] if pretrained else None
    m = refmodel.TwoTowerModel(cfg, table)
    m.load_state_dict({k: torch.tensor(v) for k, v in sd_np.items()})
    return m, refmodel


def quirk_tokens(rng, B, T, V):
    """Rows exercising SURVEY quirk #1: zeros in the middle / at the front."""
    x = np.zeros((B, T), dtype=np.int64)
    lens = rng.integers(1, T + 1, size=B)
    lens[0] = T
    for b in range(B):
        x[b, :lens[b]] = rng.integers(1, V, size=lens[b])
    if B > 1 and lens[1] >= 3:
        x[1, 1] = 0            # mid-sequence "the": drops the trailing token
    if B > 2:
        x[2, :] = 0
        x[2, 0] = 0
        x[2, 1:4] = rng.integers(1, V, size=3)   # leading 0 then 3 tokens -> first 3 positions
    return x


def case_small(name, cfg, pretrained, seed):
    rng = np.random.default_rng(seed)
    sd = synth.make_state_dict(cfg, seed=seed, table_seed=seed + 100, zero_pad_row=not pretrained)
    m, refmodel = ref_model(cfg, sd, pretrained)
    m.eval()
    B, T = 7, 10
    q = quirk_tokens(rng, B, 6, cfg["VOCAB_SIZE"])
    p = quirk_tokens(rng, B, T, cfg["VOCAB_SIZE"])
    n = quirk_tokens(rng, B, T + 3, cfg["VOCAB_SIZE"])
    with torch.no_grad():
        qe = m.encode_query(torch.tensor(q))
        pe = m.encode_document(torch.tensor(p))
        ne = m.encode_document(torch.tensor(n))
        loss = refmodel.triplet_loss_cosine((qe, pe, ne), margin=cfg["MARGIN"])
    # one live training step (main.py:244-259) with dropout disabled by the case config
    m.train()
    opt = torch.optim.Adam(m.parameters(), lr=cfg["LR"])
    opt.zero_grad()
    l2 = refmodel.triplet_loss_cosine((m.encode_query(torch.tensor(q)), m.encode_document(torch.tensor(p)),
                                       m.encode_document(torch.tensor(n))), margin=cfg["MARGIN"])
    l2.backward()
    grads = {k: (v.grad.detach().numpy().copy() if v.grad is not None else np.zeros(v.shape, np.float32))
             for k, v in m.named_parameters() if v.requires_grad}
    gnorm = float(torch.nn.utils.clip_grad_norm_(m.parameters(), max_norm=1.0))
    opt.step()
    after = {k: v.detach().numpy().copy() for k, v in m.state_dict().items()}
    blob = {"cfg": json.dumps(cfg), "pretrained": pretrained, "q": q, "p": p, "n": n,
            "q_emb": qe.numpy(), "p_emb": pe.numpy(), "n_emb": ne.numpy(), "loss": float(loss),
            "train_loss": float(l2), "grad_norm": gnorm}
    blob.update({f"w::{k}": v for k, v in sd.items()})
    blob.update({f"g::{k}": v for k, v in grads.items()})
    blob.update({f"a::{k}": v for k, v in after.items() if not k.endswith("embedding.weight") or not pretrained})
    np.savez_compressed(OUT / f"{name}.npz", **blob)
    print(name, "loss", float(loss), "gnorm", gnorm)


def case_cfgdims():
    """config.json dims (H=256, 2 layers, bidirectional, E=200) on a 2000-word vocabulary;
    weights are re-creatable from synth.make_state_dict(seed=0) so only outputs are stored."""
    cfg = synth.default_config(vocab_size=2000, embed_dim=200)
    cfg["DROPOUT"] = 0.0
    sd = synth.make_state_dict(cfg, seed=0, table_seed=1)
    m, refmodel = ref_model(cfg, sd, True)
    m.eval()
    q, _ = synth.make_tokens(16, "query", 2000, seed=11)
    p, _ = synth.make_tokens(16, "passage", 2000, seed=12, lengths=np.random.default_rng(12).integers(8, 48, 16))
    n, _ = synth.make_tokens(16, "passage", 2000, seed=13, lengths=np.random.default_rng(13).integers(8, 48, 16))
    with torch.no_grad():
        qe, pe, ne = m.encode_query(torch.tensor(q)), m.encode_document(torch.tensor(p)), m.encode_document(torch.tensor(n))
        loss = refmodel.triplet_loss_cosine((qe, pe, ne), margin=0.5)
    m.train()
    opt = torch.optim.Adam(m.parameters(), lr=cfg["LR"])
    losses, gnorms = [], []
    for _ in range(2):
        opt.zero_grad()
        l2 = refmodel.triplet_loss_cosine((m.encode_query(torch.tensor(q)), m.encode_document(torch.tensor(p)),
                                           m.encode_document(torch.tensor(n))), margin=0.5)
        l2.backward()
        if not gnorms:
            grad_l2 = {k: float(v.grad.norm()) for k, v in m.named_parameters() if v.requires_grad}
            grad_head = {k: v.grad.detach().reshape(-1)[:64].numpy().copy()
                         for k, v in m.named_parameters() if v.requires_grad}
        gnorms.append(float(torch.nn.utils.clip_grad_norm_(m.parameters(), max_norm=1.0)))
        opt.step()
        losses.append(float(l2))
    after_head = {k: v.detach().reshape(-1)[:64].numpy().copy() for k, v in m.state_dict().items()
                  if not k.endswith("embedding.weight")}
    blob = {"cfg": json.dumps(cfg), "q": q, "p": p, "n": n, "q_emb": qe.numpy(), "p_emb": pe.numpy(),
            "n_emb": ne.numpy(), "loss": float(loss), "train_losses": np.array(losses),
            "grad_norms": np.array(gnorms), "grad_l2": json.dumps(grad_l2)}
    blob.update({f"gh::{k}": v for k, v in grad_head.items()})
    blob.update({f"ah::{k}": v for k, v in after_head.items()})
    np.savez_compressed(OUT / "cfgdims.npz", **blob)
    print("cfgdims loss", float(loss), "train", losses, "gnorm", gnorms)


def case_search():
    """`torch.matmul` + `torch.topk` (evaluators.py:185-186) on seeded unit rows."""
    D = synth.make_unit_rows(20000, 256, seed=3)
    Q = synth.make_unit_rows(5, 256, seed=4)
    vals, idx = torch.topk(torch.matmul(torch.tensor(Q), torch.tensor(D).t()), k=50, dim=1)
    np.savez_compressed(OUT / "search.npz", n_docs=20000, dim=256, doc_seed=3, query_seed=4,
                        scores=vals.numpy(), idx=idx.numpy())
    print("search ok", vals[0, :3])


def case_inferencer_hybrid():
    """Run the reference `QueryInferencer` and `SimpleHybridRetriever` end to end on a
    synthetic artefact directory (query_inferencer.py:23-75, simple_hybrid.py:16-66)."""
    words = ["the", "machine", "learning", "deep", "neural", "network", "data", "text", "image", "video",
             "language", "natural", "vision", "computer", "model", "layer", "search", "query", "document",
             "retrieval", "vector", "index", "tower", "embedding", "train", "loss", "cosine", "score",
             "rank", "passage", ".", ",", "?"]
    w2i = {w: i for i, w in enumerate(words)}
    cfg = {"HIDDEN_DIM": 16, "RNN_TYPE": "GRU", "NUM_LAYERS": 2, "BIDIRECTIONAL": True, "DROPOUT": 0.2,
           "MARGIN": 0.5, "NORMALIZE_OUTPUT": True, "EMBED_DIM": 12}
    rng = np.random.default_rng(77)
    docs = [" ".join(rng.choice(words[1:30], size=int(rng.integers(4, 14)))) + " ." for _ in range(40)]
    queries = ["machine learning model", "deep neural network layer", "unknownword zzz", "the the", "vector search index ?"]
    with tempfile.TemporaryDirectory() as td:
        td = Path(td)
        art = td / "artifacts" / "run-x"
        art.mkdir(parents=True)
        (td / "frontend").mkdir()
        (td / "frontend" / "config.json").write_text(json.dumps({"ARTIFACTS_PATH": str(art)}))
        with open(art / "word_to_idx.pkl", "wb") as f:
            pickle.dump(w2i, f)
        full = dict(cfg, VOCAB_SIZE=len(w2i) + 1)       # tokenizer appends <UNK> (tokenizer.py:20-24)
        sd = synth.make_state_dict(full, seed=21, table_seed=22)
        torch.save({k: torch.tensor(v) for k, v in sd.items()}, art / "model.pth")
        (art / "config.json").write_text(json.dumps(cfg))
        cwd = os.getcwd()
        os.chdir(td)
        try:
            import simple_hybrid as ref_sh           # imports query_inferencer at CWD
            r = ref_sh.SimpleHybridRetriever(str(art), alpha=0.6)
            r.dense_retriever.device = torch.device("cpu")
            r.dense_retriever.model.to("cpu")
            r.fit(docs)
            embs, res_idx, res_score, raises = [], [], [], []
            for q in queries:
                try:                      # "the the" -> ids [0, 0] -> length 0 -> RuntimeError (quirk #2)
                    e = r.dense_retriever.get_query_embedding(q)
                    out = r.search(q, top_k=10)
                    raises.append(False)
                except RuntimeError:
                    e, out = np.zeros(cfg["HIDDEN_DIM"], np.float32), []
                    raises.append(True)
                embs.append(e)
                res_idx.append([docs.index(d) for d, _ in out] or [-1] * 10)
                res_score.append([float(s) for _, s in out] or [0.0] * 10)
            embs = np.stack(embs)
        finally:
            os.chdir(cwd)
    blob = {"cfg": json.dumps(full), "words": json.dumps(words), "docs": json.dumps(docs),
            "queries": json.dumps(queries), "alpha": 0.6, "query_emb": embs,
            "doc_emb": r.doc_embeddings, "raises": np.array(raises), "res_idx": np.array(res_idx), "res_score": np.array(res_score)}
    blob.update({f"w::{k}": v for k, v in sd.items()})
    np.savez_compressed(OUT / "inferencer_hybrid.npz", **blob)
    print("inferencer/hybrid ok", res_idx[0][:5])


def case_artifacts_eval():
    """Run the reference artefact writer (`backend/main.py:92-153`), `BatchEvaluator` and
    `CorpusEvaluator` (`backend/evaluators.py:9-209`) on a synthetic triplet set at H = 256, and the
    reference `QueryInferencer` on the directory the writer produced.  `backend/main.py` imports
    `data_loader` -> `fastparquet` (absent here): an empty stub module named `fastparquet` is put into
    `sys.modules` for the import; none of the functions exercised touches it."""
    import random
    import types
    sys.modules.setdefault("fastparquet", types.ModuleType("fastparquet"))
    os.environ.setdefault("WANDB_MODE", "disabled")
    import main as ref_main                      # /root/reference/backend/main.py, unmodified
    import evaluators as ref_eval
    words = ["the", "machine", "learning", "deep", "neural", "network", "data", "text", "image", "video",
             "language", "natural", "vision", "computer", "model", "layer", "search", "query", "document",
             "retrieval", "vector", "index", "tower", "embedding", "train", "loss", "cosine", "score",
             "rank", "passage", "and", "of", "in", ".", ",", "?"]
    w2i = {w: i for i, w in enumerate(words)}
    cfg = {"HIDDEN_DIM": 256, "RNN_TYPE": "GRU", "NUM_LAYERS": 1, "BIDIRECTIONAL": True, "DROPOUT": 0.0,
           "MARGIN": 0.5, "NORMALIZE_OUTPUT": True, "EMBED_DIM": 12, "BATCH_SIZE": 64}
    rng = np.random.default_rng(91)
    pool = [" ".join(rng.choice(words[1:33], size=int(rng.integers(5, 16)))) + " ." for _ in range(90)]
    queries = [" ".join(rng.choice(words[1:30], size=int(rng.integers(2, 6)))) for _ in range(36)]
    triplets = []
    pos_order = rng.permutation(90)                       # every positive is used once: BatchEvaluator's rank of a
    for qi, q in enumerate(queries):                      # duplicated positive would hinge on torch.sort's tie order
        for _ in range(int(rng.integers(1, 3))):          # several positives per query
            triplets.append((q, pool[int(pos_order[len(triplets)])], pool[int(rng.integers(0, 90))]))
    val = triplets[: len(triplets) // 2]
    with tempfile.TemporaryDirectory() as td:
        td = Path(td)
        art = td / "artifacts" / "run-y"
        (td / "frontend").mkdir()
        (td / "frontend" / "config.json").write_text(json.dumps({"ARTIFACTS_PATH": str(art)}))
        w2i_path = td / "word_to_idx.pkl"
        with open(w2i_path, "wb") as f:
            pickle.dump(w2i, f)
        import tokenizer as ref_tok
        tok = ref_tok.PretrainedTokenizer(str(w2i_path))
        full = dict(cfg, VOCAB_SIZE=tok.vocab_size())
        sd = synth.make_state_dict(full, seed=31, table_seed=32)
        model, _ = ref_model(full, sd, True)
        model.device = torch.device("cpu")
        run_cfg = dict(cfg, WORD_TO_IDX_PATH=str(w2i_path))
        ref_main.save_inference_artifacts(art, model, run_cfg, tok, {"train": triplets, "validation": val})
        with open(art / "documents.pkl", "rb") as f:
            documents = pickle.load(f)
        doc_emb = np.load(art / "document_embeddings.npy")
        with open(art / "tfidf_artifacts.pkl", "rb") as f:
            tf = pickle.load(f)
        mat = tf["matrix"].tocsr()
        mat.sort_indices()
        vocab = tf["vectorizer"].vocabulary_
        saved_cfg = json.loads((art / "config.json").read_text())
        state_keys = list(torch.load(art / "model.pth").keys())
        # evaluators
        loader = torch.utils.data.DataLoader(ref_main.TripletDataset(val, tok), batch_size=16, shuffle=False,
                                             collate_fn=ref_main.collate_fn)
        bm, bl = ref_eval.BatchEvaluator(top_k=[1, 5, 10]).evaluate(model, loader, torch.device("cpu"), full)
        random.seed(1234)
        # defaults (max_candidates=1000, max_queries=50) exceed the set sizes: nothing is sub-sampled, so the
        # result does not depend on the per-process string hash order of `list(set(...))`
        cm = ref_eval.CorpusEvaluator(top_k=[1, 5, 10]).evaluate(
            model, val, tok, torch.device("cpu"))
        # the reference reader on the directory the reference writer produced
        cwd = os.getcwd()
        os.chdir(td)
        try:
            import query_inferencer as ref_qi
            inf = ref_qi.QueryInferencer(str(art), device=torch.device("cpu"))
            probe = ["machine learning model", "deep neural network layer", "vector search index ?", "zzz qqq"]
            q_emb = np.stack([inf.get_query_embedding(q) for q in probe])
        finally:
            os.chdir(cwd)
    blob = {"cfg": json.dumps(full), "saved_cfg": json.dumps(saved_cfg), "words": json.dumps(words),
            "triplets": json.dumps(triplets), "n_val": len(val), "documents": json.dumps(documents),
            "doc_emb": doc_emb, "tfidf_indptr": mat.indptr, "tfidf_indices": mat.indices, "tfidf_data": mat.data,
            "tfidf_vocab": json.dumps({k: int(v) for k, v in vocab.items()}), "tfidf_shape": np.array(mat.shape),
            "state_keys": json.dumps(state_keys), "weight_seeds": np.array([31, 32]),
            "batch_metrics": json.dumps({k: float(v) for k, v in bm.items()}), "batch_loss": float(bl),
            "corpus_metrics": json.dumps({k: float(v) for k, v in cm.items()}), "corpus_seed": 1234,
            "probe_queries": json.dumps(probe), "probe_emb": q_emb}
    np.savez_compressed(OUT / "artifacts_eval.npz", **blob)
    print("artifacts/eval ok", bm, cm)


def main():
    torch.manual_seed(0)
    torch.set_num_threads(1)          # deterministic reductions
    OUT.mkdir(parents=True, exist_ok=True)
    base = {"RNN_TYPE": "GRU", "DROPOUT": 0.0, "LR": 1e-3, "MARGIN": 0.5}
    case_small("small_bi2", dict(base, VOCAB_SIZE=64, EMBED_DIM=12, HIDDEN_DIM=16, NUM_LAYERS=2,
                                 BIDIRECTIONAL=True, NORMALIZE_OUTPUT=True), True, 5)
    case_small("small_uni1", dict(base, VOCAB_SIZE=40, EMBED_DIM=8, HIDDEN_DIM=16, NUM_LAYERS=1,
                                  BIDIRECTIONAL=False, NORMALIZE_OUTPUT=False), True, 6)
    case_small("small_bi1_trainable_table", dict(base, VOCAB_SIZE=48, EMBED_DIM=8, HIDDEN_DIM=8, NUM_LAYERS=1,
                                                 BIDIRECTIONAL=True, NORMALIZE_OUTPUT=True), False, 7)
    case_small("small_uni2", dict(base, VOCAB_SIZE=40, EMBED_DIM=20, HIDDEN_DIM=32, NUM_LAYERS=2,
                                  BIDIRECTIONAL=False, NORMALIZE_OUTPUT=True), True, 8)
    case_cfgdims()
    case_search()
    case_inferencer_hybrid()
    case_artifacts_eval()


if __name__ == "__main__":
    main()
