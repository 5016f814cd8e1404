"""Drop-in `BatchEvaluator` / `CorpusEvaluator` — reference `backend/evaluators.py:9-209` — on the device.

Same classes, constructor arguments, `evaluate(...)` signatures and metric names.  What changes is
where the work happens:

* `BatchEvaluator`: the reference builds the full [Q, Q] similarity matrix and sorts every row to
  find the positive's rank (`evaluators.py:49-73`); here one streaming kernel (`ttr_positive_rank`)
  counts the documents that beat the positive — the rank a stable descending sort gives — and the
  metrics are reductions over that rank vector.
* `CorpusEvaluator`: documents are encoded in length-bucketed batches (`encode.encode_rows`) instead
  of 64 at a time, all sampled queries are encoded in one batch, and `matmul` + `topk`
  (`evaluators.py:185-186`) is the fused exact top-k kernel.  Sampling uses Python's `random`
  exactly like the reference, so a seeded run draws the same candidates and queries.
"""
from __future__ import annotations

import random
from typing import Dict, List, Tuple

import numpy as np
import torch

from . import _lib
from .encode import encode_rows
from .index import search_topk
from .model import triplet_loss_cosine


def positive_ranks(query_embs: torch.Tensor, doc_embs: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """int32 [Q]: 1-based rank of document target[i] for query i (ties: lower index first)."""
    _lib.require_cuda(query_embs, "positive_ranks(query_embs)")
    q = query_embs.detach().contiguous().float()
    d = doc_embs.detach().contiguous().float()
    t = target.to(device=q.device, dtype=torch.int64).contiguous()
    out = torch.empty(q.shape[0], dtype=torch.int32, device=q.device)
    _lib.call("ttr_positive_rank", q, d, t, q.shape[0], d.shape[0], q.shape[1], out, None)
    return out


class BatchEvaluator:
    """`evaluators.py:9-80`: 1:1 query/positive mapping over the concatenated validation batches."""

    def __init__(self, top_k: List[int] = [1, 5, 10]):
        self.top_k = top_k

    def evaluate(self, model, val_loader, device: torch.device, config: Dict):
        model.eval()
        all_q, all_d = [], []
        total = torch.zeros((), dtype=torch.float32, device=device)
        n_batches = 0
        with torch.no_grad():
            for queries, pos_docs, neg_docs in val_loader:
                queries, pos_docs, neg_docs = queries.to(device), pos_docs.to(device), neg_docs.to(device)
                q = model.encode_query(queries)
                p = model.encode_document(pos_docs)
                n = model.encode_document(neg_docs)
                total += triplet_loss_cosine((q, p, n), margin=config.get("MARGIN", 0.2))   # no per-batch .item() sync
                n_batches += 1
                all_q.append(q)
                all_d.append(p)
        if not all_q:
            return {}, 0
        query_embs, doc_embs = torch.cat(all_q), torch.cat(all_d)
        nq = query_embs.shape[0]
        ranks = positive_ranks(query_embs, doc_embs, torch.arange(nq, device=device)).double()
        stats = torch.stack([(ranks <= k).double().mean() for k in self.top_k] + [(1.0 / ranks).mean()]).cpu().tolist()
        metrics = {f"Recall@{k}": v for k, v in zip(self.top_k, stats)}
        metrics["MRR"] = stats[-1]
        return metrics, float(total) / n_batches


class CorpusEvaluator:
    """`evaluators.py:83-209`: unique queries against a sampled candidate pool, several positives per query."""

    def __init__(self, top_k: List[int] = [1, 5, 10], max_candidates: int = 1000, max_queries: int = 50):
        self.top_k = top_k
        self.max_candidates = max_candidates
        self.max_queries = max_queries

    def evaluate(self, model, val_data: List[Tuple[str, str, str]], tokenizer, device: torch.device):
        model.eval()
        query_to_positives: Dict[str, set] = {}
        all_docs = set()
        for query, pos_doc, neg_doc in val_data:
            query_to_positives.setdefault(query, set()).add(pos_doc)
            all_docs.add(pos_doc)
            all_docs.add(neg_doc)
        unique_queries = list(query_to_positives.keys())
        unique_docs = list(all_docs)
        if len(unique_docs) > self.max_candidates:
            unique_docs = random.sample(unique_docs, self.max_candidates)
        doc_embeddings = self._compute_document_embeddings(model, unique_docs, tokenizer, device)
        sample_queries = random.sample(unique_queries, min(self.max_queries, len(unique_queries)))
        metrics = {f"Recall@{k}": [] for k in self.top_k}
        metrics.update({f"Hit@{k}": [] for k in self.top_k})
        if sample_queries:
            q_emb = encode_rows(model.query_encoder, [tokenizer.encode(q) for q in sample_queries], device)
            kmax = max(self.top_k)
            if kmax > doc_embeddings.shape[0]:
                raise RuntimeError("selected index k out of range")          # what torch.topk raises (evaluators.py:186)
            if q_emb.shape[1] == 256 and kmax <= 64:
                _, top = search_topk(q_emb, doc_embeddings, kmax)
            else:
                _, top = torch.topk(q_emb @ doc_embeddings.t(), k=kmax, dim=1)
            top = top.cpu().numpy()
            doc_set = set(unique_docs)
            for qi, query in enumerate(sample_queries):
                known = query_to_positives[query]
                available = [d for d in known if d in doc_set]
                if not available:
                    continue
                for k in self.top_k:
                    top_k_docs = [unique_docs[i] for i in top[qi, :k]]
                    found = len([d for d in top_k_docs if d in known])
                    metrics[f"Recall@{k}"].append(found / len(available))
                    metrics[f"Hit@{k}"].append(1 if found > 0 else 0)
        return {name: (float(np.mean(v)) if v else 0.0) for name, v in metrics.items()}

    def _compute_document_embeddings(self, model, documents: List[str], tokenizer, device: torch.device):
        return encode_rows(model.doc_encoder, [tokenizer.encode(d) for d in documents], device)
