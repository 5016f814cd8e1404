"""Artefact directory round trip — reference writer `backend/main.py:92-153`, readers
`backend/query_inferencer.py:36-56` and `frontend/main.py:54-69`.

File set and formats are the reference's, so a directory written here loads in the reference and the
other way round:
    model.pth                 torch state_dict, reference key names
    config.json               training config + VOCAB_SIZE + EMBED_DIM
    word_to_idx.pkl           tokenizer vocabulary
    documents.pkl             list[str], row i <-> embedding row i
    document_embeddings.npy   fp32 [N, H] C-contiguous
    tfidf_artifacts.pkl       {'vectorizer': TfidfVectorizer, 'matrix': CSR float64 [N, F]}
The document embeddings come from the bulk encoder (length-bucketed batches, no per-batch D2H)
instead of the reference's 64-string loop; `load_index` puts a row shard of the matrix and of the
TF-IDF CSR straight into HBM.
"""
from __future__ import annotations

import json
import pickle
import shutil
from pathlib import Path
from typing import Dict, Optional

import numpy as np
import torch

from .encode import encode_rows
from .index import CsrF64, ShardedIndex, shard_bounds


def save_inference_artifacts(output_dir, model, config: Dict, tokenizer, datasets: Dict, max_features: int = 20000):
    """`backend/main.py:92-153` with the same arguments (`datasets`: split -> list of (q, pos, neg) strings)."""
    from sklearn.feature_extraction.text import TfidfVectorizer
    output_dir = Path(output_dir)
    output_dir.mkdir(parents=True, exist_ok=True)
    torch.save({k: v.detach().cpu() for k, v in model.state_dict().items()}, output_dir / "model.pth")
    cfg = dict(config)
    cfg["VOCAB_SIZE"] = tokenizer.vocab_size()
    cfg["EMBED_DIM"] = model.query_encoder.embedding.embedding_dim
    with open(output_dir / "config.json", "w") as fh:
        json.dump(cfg, fh, indent=4)
    src = config.get("WORD_TO_IDX_PATH")
    if src is not None and Path(src).resolve() != (output_dir / "word_to_idx.pkl").resolve():
        shutil.copyfile(src, output_dir / "word_to_idx.pkl")
    all_docs = set()
    for split in datasets.values():
        for _, pos_doc, neg_doc in split:
            all_docs.add(pos_doc)
            all_docs.add(neg_doc)
    unique_docs = list(all_docs)
    device = next(p for p in model.parameters()).device
    emb = encode_rows(model.doc_encoder, [tokenizer.encode(d) for d in unique_docs], device)
    with open(output_dir / "documents.pkl", "wb") as fh:
        pickle.dump(unique_docs, fh)
    np.save(output_dir / "document_embeddings.npy", np.ascontiguousarray(emb.cpu().numpy()))
    vec = TfidfVectorizer(stop_words="english", max_features=max_features)
    mat = vec.fit_transform(unique_docs)
    with open(output_dir / "tfidf_artifacts.pkl", "wb") as fh:
        pickle.dump({"vectorizer": vec, "matrix": mat}, fh)
    return unique_docs, emb


def load_corpus_artifacts(artifacts_path):
    """(documents list, embeddings np.float32 [N, H] memory-mapped, vectorizer, tfidf CSR) — raises
    FileNotFoundError for a missing file like the reference's `open` calls."""
    art = Path(artifacts_path)
    with open(art / "tfidf_artifacts.pkl", "rb") as fh:
        tf = pickle.load(fh)
    with open(art / "documents.pkl", "rb") as fh:
        docs = pickle.load(fh)
    emb = np.load(art / "document_embeddings.npy", mmap_mode="r")
    if emb.shape[0] != len(docs) or tf["matrix"].shape[0] != len(docs):
        raise ValueError(f"artefacts disagree: {len(docs)} documents, {emb.shape[0]} embeddings, "
                         f"{tf['matrix'].shape[0]} TF-IDF rows")
    return docs, emb, tf["vectorizer"], tf["matrix"].tocsr()


def load_index(artifacts_path, device, group=None, world: Optional[int] = None, rank: Optional[int] = None):
    """Row shard [r*ceil(N/R), ...) of `document_embeddings.npy` and of the TF-IDF matrix -> ShardedIndex."""
    import torch.distributed as dist
    docs, emb, vec, mat = load_corpus_artifacts(artifacts_path)
    if world is None:
        world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        rank = dist.get_rank(group) if world > 1 else 0
    lo, hi = shard_bounds(len(docs), world, rank)
    local = torch.from_numpy(np.array(emb[lo:hi], dtype=np.float32, order="C")).to(device)
    csr = CsrF64.from_scipy(mat[lo:hi], device, row_offset=lo)
    return ShardedIndex(local, lo, len(docs), group=group, tfidf_local=csr), docs, vec, mat
