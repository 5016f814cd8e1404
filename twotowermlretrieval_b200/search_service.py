"""The `/search` handler of the reference frontend (`frontend/main.py:102-210`) on the GPU index.

`SearchService.search(query, alpha)` returns the reference's response body
    {"query", "alpha", "results": [{"rank", "id", "doc", "score", "dense_score", "tfidf_score"}, ...]}
with the same three branches:
  * alpha == 0      corpus-wide TF-IDF cosine, top 10, hits with score <= 1e-5 dropped (`:119-147`)
  * otherwise       top-50 by dense similarity, `dense_score = 1 - dist` (Chroma's default squared-L2
                    space on unit vectors: 2 cos - 1), TF-IDF cosine of the 50 candidates (zeros when
                    the query has no vocabulary hit, `:169-175`), `alpha` blend, stable sort, top 10
The Chroma ANN lookup (`collection.query`, `:153-156`) is replaced by the exact fused score + top-k
kernel over the resident document matrix; candidate TF-IDF rows are the stored `doc_tfidf_matrix`
rows (what re-transforming the candidate strings yields).  The HTTP layer itself (FastAPI, CORS, HTML)
is out of scope; `search` takes the two fields of `QueryInput` directly.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from . import _lib
from .artifacts import load_index
from .index import CsrF64, _SEARCH_LOCK
from .query_inferencer import QueryInferencer

N_CANDIDATES = 50      # frontend/main.py:155
N_RESULTS = 10         # frontend/main.py:196


class SearchService:
    def __init__(self, artifacts_path: str, device: Optional[torch.device] = None, group=None, space="l2"):
        self.inferencer = QueryInferencer(artifacts_path, device=device)
        self.device = self.inferencer.device
        self.index, self.documents, self.vectorizer, self.doc_tfidf_matrix = load_index(artifacts_path, self.device,
                                                                                       group=group)
        self.space = space
        self._ws = None

    def _query_csr(self, query: str):
        q = self.vectorizer.transform([query]).tocsr()
        q.sort_indices()
        return q

    def search(self, query: str, alpha: float = 0.5) -> dict:
        q_row = self._query_csr(query)
        if alpha == 0.0:
            top = self._keyword(q_row)
        else:
            top = self._hybrid(query, q_row, alpha)
        return {"query": query, "alpha": alpha,
                "results": [{"rank": i + 1, "id": f"result-{i + 1}", **res} for i, res in enumerate(top)]}

    # alpha == 0: corpus-wide keyword search (frontend/main.py:119-147)
    def _keyword(self, q_row):
        if self.index.world != 1:
            raise NotImplementedError("keyword branch is served from a single-GPU index")
        docs = self.index.docs
        N, D = docs.shape
        k = min(N_RESULTS, N)
        dev = self.device
        q_idx = torch.as_tensor(q_row.indices.astype(np.int32), device=dev)
        q_val = torch.as_tensor(q_row.data.astype(np.float64), device=dev)
        lib = _lib.load()
        nbytes = lib.ttr_blend_topk_workspace_bytes(k)
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = torch.empty(int(nbytes), dtype=torch.uint8, device=dev)
        out_s = torch.empty(k, dtype=torch.float64, device=dev)
        out_i = torch.empty(k, dtype=torch.int64, device=dev)
        zero_q = torch.zeros(D, dtype=torch.float32, device=dev)
        csr = self.index.tfidf
        with _SEARCH_LOCK:                  # the per-object workspace is shared by concurrent callers
            _lib.call("ttr_blend_topk", zero_q, 0.0, docs, N, D, csr.indptr, csr.indices, csr.data, q_idx, q_val,
                  int(q_idx.numel()), 0.0, k, out_s, out_i, None, self._ws)
        res = []
        for i, s in zip(out_i.cpu().tolist(), out_s.cpu().tolist()):
            if i >= 0 and s > 1e-5:
                res.append({"doc": self.documents[i], "score": float(s), "dense_score": 0.0, "tfidf_score": float(s)})
        return res

    # hybrid: dense top-50 -> TF-IDF of the candidates -> blend -> top-10 (frontend/main.py:149-198)
    def _hybrid(self, query: str, q_row, alpha: float):
        q_emb = torch.from_numpy(self.inferencer.get_query_embedding(query)).to(self.device).unsqueeze(0)
        q_csr = CsrF64.from_arrays(q_row.indptr, q_row.indices, q_row.data, self.device)
        kc = min(N_CANDIDATES, self.index.n_total)
        out = self.index.search_hybrid(q_emb, q_csr, alpha, k=kc, top_n=min(N_RESULTS, kc), space=self.space)
        idx = out["idx"][0].cpu().tolist()
        fin, sem, tf = (out[k][0].cpu().tolist() for k in ("final", "semantic", "tfidf"))
        return [{"doc": self.documents[i], "score": float(f), "dense_score": float(s), "tfidf_score": float(t)}
                for i, f, s, t in zip(idx, fin, sem, tf) if i >= 0]
