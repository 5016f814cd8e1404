"""Drop-in `SimpleHybridRetriever` — reference `backend/simple_hybrid.py:13-66`.

`fit` keeps the reference's choices (documents go through the QUERY tower, quirk #10;
`TfidfVectorizer(stop_words='english', max_features=10000)`), but encodes all documents in
length-bucketed batches instead of one forward per document.  `search` blends
alpha*dense_cos + (1-alpha)*tfidf_cos over the whole corpus and returns the `top_k` best
`(document, score)` pairs in `np.argsort(...)[::-1]` order, like `simple_hybrid.py:57-66`.
The text vectoriser stays host-side sklearn (string processing is out of scope).
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np
import torch

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
        self._doc_norm = None

    def fit(self, documents: List[str]):
        self.documents = list(documents)
        self.tfidf_matrix = self.tfidf.fit_transform(self.documents)
        self._doc_dev = self.dense_retriever.encode_queries(self.documents)     # query tower, quirk #10
        self.doc_embeddings = self._doc_dev.cpu().numpy()
        self._doc_norm = torch.linalg.vector_norm(self._doc_dev, dim=1)

    def search(self, query: str, top_k: int = 10) -> List[Tuple[str, float]]:
        dev = self._doc_dev.device
        q_tfidf = self.tfidf.transform([query])
        tfidf_scores = np.asarray((self.tfidf_matrix @ q_tfidf.T).todense()).ravel()   # L2 rows: cosine == dot
        q = torch.from_numpy(self.dense_retriever.get_query_embedding(query)).to(dev)
        # sklearn cosine_similarity (simple_hybrid.py:53-54): normalise both sides (zero rows stay zero)
        qn = torch.linalg.vector_norm(q)
        dense = (self._doc_dev @ q) / (self._doc_norm * qn).clamp_min(torch.finfo(torch.float32).tiny)
        dense = torch.where((self._doc_norm == 0) | (qn == 0), torch.zeros_like(dense), dense)
        combined = self.alpha * dense.double().cpu().numpy() + (1 - self.alpha) * tfidf_scores
        top = np.argsort(combined)[::-1][:top_k]
        return [(self.documents[i], combined[i]) for i in top]
