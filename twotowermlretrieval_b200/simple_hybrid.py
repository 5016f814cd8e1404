"""Drop-in `SimpleHybridRetriever` — reference `backend/simple_hybrid.py:13-66`.

`fit` keeps the reference's choices (documents go through the QUERY tower, quirk #10;
`TfidfVectorizer(stop_words='english', max_features=10000)`), but encodes all documents in
length-bucketed batches instead of one forward per document, and keeps the embeddings and the
TF-IDF CSR resident on the device.  `search` runs one fused kernel pass over the corpus
(`ttr_blend_topk`): alpha*dense_cos + (1-alpha)*tfidf_cos for every document and the `top_k`
best `(document, score)` pairs in `np.argsort(...)[::-1]` order, like `simple_hybrid.py:57-66`.
The text vectoriser stays host-side sklearn (string processing is out of scope).
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np
import torch

from .index import KMAX, CsrF64, blend_topk
from .query_inferencer import QueryInferencer


class SimpleHybridRetriever:
    def __init__(self, artifacts_path: str, alpha: float = 0.5, device=None):
        from sklearn.feature_extraction.text import TfidfVectorizer
        self.dense_retriever = QueryInferencer(artifacts_path, device=device)
        self.alpha = alpha
        self.tfidf = TfidfVectorizer(stop_words="english", max_features=10000)
        self.documents: List[str] = []
        self.doc_embeddings = None          # np.float32 [N, H], like the reference attribute
        self._doc_dev = None                # resident copy, fp32 [N, H]
        self._csr = None                    # resident TF-IDF matrix

    def fit(self, documents: List[str]):
        self.documents = list(documents)
        self.tfidf_matrix = self.tfidf.fit_transform(self.documents)
        dev = self.dense_retriever.device
        self._doc_dev = self.dense_retriever.encode_queries(self.documents).contiguous()   # query tower, quirk #10
        self.doc_embeddings = self._doc_dev.cpu().numpy()
        self._csr = CsrF64.from_scipy(self.tfidf_matrix, dev)

    def search(self, query: str, top_k: int = 10) -> List[Tuple[str, float]]:
        dev = self._doc_dev.device
        N, D = self._doc_dev.shape
        k = min(top_k, N)
        if not 1 <= k <= KMAX:
            raise ValueError(f"SimpleHybridRetriever.search: top_k={top_k} outside [1, {KMAX}] — the fused blend + "
                             "top-k kernel keeps at most 64 results per query (the reference default is 10)")
        q_row = self.tfidf.transform([query]).tocsr()
        q_row.sort_indices()
        q_idx = torch.as_tensor(q_row.indices.astype(np.int32), device=dev)
        q_val = torch.as_tensor(q_row.data.astype(np.float64), device=dev)
        q_np = self.dense_retriever.get_query_embedding(query)
        q = torch.from_numpy(q_np).to(dev)
        out_s, out_i = blend_topk(q, float(np.linalg.norm(q_np)), self._doc_dev, self._csr, q_idx, q_val,
                                  self.alpha, k)
        idx, sc = out_i.cpu().tolist(), out_s.cpu().tolist()
        return [(self.documents[i], s) for i, s in zip(idx, sc) if i >= 0]
