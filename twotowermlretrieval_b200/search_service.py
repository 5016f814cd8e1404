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

from .artifacts import load_index
from .index import CsrF64, blend_topk
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
        """Every rank scans the TF-IDF rows it owns (blend kernel with alpha = 0) for its 10 best; on a sharded index
        the [10] lists (fp64 score, global row id) are all-gathered and the 10 best of the union are kept, ordered by
        (score desc, id desc) — identical on every rank and to the single-shard result."""
        index = self.index
        docs = index.docs
        n_local, D = docs.shape
        k = min(N_RESULTS, index.n_total)
        dev = self.device
        q_idx = torch.as_tensor(q_row.indices.astype(np.int32), device=dev)
        q_val = torch.as_tensor(q_row.data.astype(np.float64), device=dev)
        zero_q = torch.zeros(D, dtype=torch.float32, device=dev)
        k_loc = min(k, n_local)
        out_s = torch.full((k,), float("-inf"), dtype=torch.float64, device=dev)
        out_i = torch.full((k,), -1, dtype=torch.int64, device=dev)
        if k_loc > 0:
            s_loc, i_loc = blend_topk(zero_q, 0.0, docs, index.tfidf, q_idx, q_val, 0.0, k_loc)
            out_s[:k_loc] = s_loc
            out_i[:k_loc] = torch.where(i_loc >= 0, i_loc + index.row_offset, i_loc)
        if index.world > 1:
            import torch.distributed as dist
            gs = torch.empty(index.world, k, dtype=torch.float64, device=dev)
            gi = torch.empty(index.world, k, dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(gs, out_s, group=index.group)
            dist.all_gather_into_tensor(gi, out_i, group=index.group)
            gs, gi = gs.reshape(-1), gi.reshape(-1)
            gs = torch.where(gi >= 0, gs, torch.full_like(gs, float("-inf")))
            # ties: higher id first, the order the single-shard kernel emits (np.argsort(...)[::-1])
            by_id = torch.argsort(gi, descending=True, stable=True)
            by_score = torch.argsort(gs[by_id], descending=True, stable=True)[:k]
            sel = by_id[by_score]
            out_s, out_i = gs[sel], gi[sel]
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
        # |q|^2: 1 for an encoded query of a normalising model, 0 for the zero vector a token-less query gets
        # (query_inferencer.py:65-69) -> Chroma's distance is |d|^2 = 1 and dense_score = 0 there, not 2*0 - 1
        q_sq = (q_emb.double() ** 2).sum(dim=1)
        unit = bool(self.inferencer.config.get("NORMALIZE_OUTPUT", True)) and float(q_sq[0]) > 0.0
        out = self.index.search_hybrid(q_emb, q_csr, alpha, k=kc, top_n=min(N_RESULTS, kc), space=self.space,
                                       q_sqnorm=None if unit else q_sq)
        idx = out["idx"][0].cpu().tolist()
        fin, sem, tf = (out[k][0].cpu().tolist() for k in ("final", "semantic", "tfidf"))
        return [{"doc": self.documents[i], "score": float(f), "dense_score": float(s), "tfidf_score": float(t)}
                for i, f, s, t in zip(idx, fin, sem, tf) if i >= 0]
