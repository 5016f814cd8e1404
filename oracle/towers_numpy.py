"""Explicit numpy restatement of the reference hot path (test oracle; see oracle/__init__.py).

Plain loops, no torch.  `dtype=np.float64` gives the high-precision answer used for
tolerance studies, `np.float32` mimics the reference's arithmetic type.
"""
from __future__ import annotations

import numpy as np


# --------------------------------------------------------------------------- towers
def effective_lengths(x: np.ndarray) -> np.ndarray:
    """`lengths = (x != 0).sum(dim=1)` — reference `backend/model.py:52`.

    NOTE: this counts non-zero ids, and the packed sequence then keeps
    the FIRST `length` positions of the row, whatever they hold."""
    return (np.asarray(x) != 0).sum(axis=1).astype(np.int64)


def _sigmoid(v):
    return 1.0 / (1.0 + np.exp(-v))


def gru_step(x_t, h, w_ih, w_hh, b_ih, b_hh):
    """One GRU cell update, PyTorch gate order r,z,n (torch `nn.GRU` docs; reached from
    `backend/model.py:59-62`):
        r = s(W_ir x + b_ir + W_hr h + b_hr);  z = s(W_iz x + b_iz + W_hz h + b_hz)
        n = tanh(W_in x + b_in + r * (W_hn h + b_hn));  h' = (1 - z) * n + z * h"""
    H = h.shape[-1]
    gi = w_ih @ x_t + b_ih
    gh = w_hh @ h + b_hh
    r = _sigmoid(gi[:H] + gh[:H])
    z = _sigmoid(gi[H:2 * H] + gh[H:2 * H])
    n = np.tanh(gi[2 * H:] + r * gh[2 * H:])
    return (1.0 - z) * n + z * h


def gru_direction(seq, w_ih, w_hh, b_ih, b_hh, reverse: bool):
    """Run one direction of one layer over one row's valid steps; returns the per-step
    outputs in original time order [len, H] and the final state (for the reverse
    direction: the state after consuming position 0)."""
    H = w_hh.shape[1]
    h = np.zeros(H, dtype=seq.dtype)
    out = np.zeros((seq.shape[0], H), dtype=seq.dtype)
    order = range(seq.shape[0] - 1, -1, -1) if reverse else range(seq.shape[0])
    for t in order:
        h = gru_step(seq[t], h, w_ih, w_hh, b_ih, b_hh)
        out[t] = h
    return out, h


def encoder_forward(sd: dict, prefix: str, x: np.ndarray, cfg: dict, dtype=np.float64,
                    dropout_masks: "list | None" = None) -> np.ndarray:
    """`RNNEncoder.forward` — reference `backend/model.py:48-75`.

    sd: state dict (numpy) with reference key names, prefix 'query_encoder' or
    'doc_encoder'.  dropout_masks (train-mode replay): for each layer l < L-1 an array
    [B, T, dirs*H] of already-scaled keep masks (0 or 1/(1-p)) applied to layer l's
    outputs, which is where nn.GRU applies inter-layer dropout (`model.py:35`)."""
    H = cfg["HIDDEN_DIM"]
    L = cfg.get("NUM_LAYERS", 1)
    bi = bool(cfg.get("BIDIRECTIONAL", False))
    x = np.asarray(x)
    lengths = effective_lengths(x)
    if (lengths <= 0).any():
        # torch: "Length of all samples has to be greater than 0" (model.py:55-57)
        raise RuntimeError("Length of all samples has to be greater than 0, "
                           "but found an element in 'lengths' that is <= 0")
    g = lambda name: np.asarray(sd[f"{prefix}.{name}"], dtype=dtype)
    table = g("embedding.weight")
    out = np.zeros((x.shape[0], H), dtype=dtype)
    for b in range(x.shape[0]):
        n = int(lengths[b])
        seq = table[x[b, :n]]                       # model.py:49 (first n positions)
        finals = []
        for layer in range(L):
            outs, finals = [], []
            for sfx, rev in ([("", False)] + ([("_reverse", True)] if bi else [])):
                o, hn = gru_direction(seq, g(f"rnn.weight_ih_l{layer}{sfx}"),
                                      g(f"rnn.weight_hh_l{layer}{sfx}"),
                                      g(f"rnn.bias_ih_l{layer}{sfx}"),
                                      g(f"rnn.bias_hh_l{layer}{sfx}"), rev)
                outs.append(o)
                finals.append(hn)
            seq = np.concatenate(outs, axis=1)
            if dropout_masks is not None and layer < L - 1:
                seq = seq * np.asarray(dropout_masks[layer][b, :n], dtype=dtype)
        if bi:                                      # model.py:65-69
            hidden = g("projection.weight") @ np.concatenate(finals) + g("projection.bias")
        else:                                       # model.py:70-71
            hidden = finals[-1]
        if cfg.get("NORMALIZE_OUTPUT", True):       # model.py:73-74, eps 1e-12
            hidden = hidden / max(np.sqrt((hidden * hidden).sum()), 1e-12)
        out[b] = hidden
    return out


def triplet_loss_cosine(q, p, n, margin: float = 0.2, dtype=np.float64) -> float:
    """`triplet_loss_cosine` — reference `backend/model.py:109-114`;
    F.cosine_similarity clamps each norm at eps=1e-8."""
    q, p, n = (np.asarray(a, dtype=dtype) for a in (q, p, n))

    def cos(a, b):
        na = np.maximum(np.sqrt((a * a).sum(1)), 1e-8)
        nb = np.maximum(np.sqrt((b * b).sum(1)), 1e-8)
        return (a * b).sum(1) / (na * nb)

    return float(np.maximum(cos(q, n) - cos(q, p) + margin, 0.0).mean())


def batch_metrics(q, p, n) -> dict:
    """`TwoTowerTrainer.compute_batch_metrics` — reference `backend/trainer.py:38-55`."""
    q, p, n = (np.asarray(a, dtype=np.float64) for a in (q, p, n))
    pos, neg = (q * p).sum(1), (q * n).sum(1)
    return {"accuracy": float((pos > neg).mean()), "similarity_gap": float((pos - neg).mean()),
            "magnitude": float(np.sqrt((q * q).sum(1)).mean()),
            "pos_similarity": float(pos.mean()), "neg_similarity": float(neg.mean())}


# --------------------------------------------------------------------------- search
def cosine_topk(Q: np.ndarray, D: np.ndarray, k: int, dtype=np.float64, chunk: int = 262144):
    """`sim = q @ D.T; torch.topk(sim, k)` — reference `backend/evaluators.py:185-186`
    (also `:50`, `:269-272`, `trainer.py:62-65`).  Ties: higher score first, then lower
    document index (torch leaves tie order unspecified; the CUDA path uses this order)."""
    Q = np.asarray(Q, dtype=dtype)
    B, N = Q.shape[0], D.shape[0]
    k = min(k, N)
    best_s = np.full((B, 0), 0, dtype=dtype)
    best_i = np.zeros((B, 0), dtype=np.int64)
    for s0 in range(0, N, chunk):
        S = Q @ np.asarray(D[s0:s0 + chunk], dtype=dtype).T
        idx = np.broadcast_to(np.arange(s0, s0 + S.shape[1], dtype=np.int64), S.shape)
        cs = np.concatenate([best_s, S], axis=1)
        ci = np.concatenate([best_i, idx], axis=1)
        order = np.lexsort((ci, -cs), axis=1)[:, :k]
        best_s = np.take_along_axis(cs, order, axis=1)
        best_i = np.take_along_axis(ci, order, axis=1)
    return best_s, best_i


def csr_row_dot(indptr, indices, data, row: int, q_idx: np.ndarray, q_val: np.ndarray) -> float:
    """Sparse·sparse dot of one L2-normalised TF-IDF row with the query row — what
    sklearn `cosine_similarity(query_tfidf, doc_tfidfs)` reduces to for `norm='l2'` rows
    (`frontend/main.py:170-172`)."""
    lo, hi = int(indptr[row]), int(indptr[row + 1])
    acc = 0.0
    for c, v in zip(indices[lo:hi], data[lo:hi]):
        hit = np.nonzero(q_idx == c)[0]
        if hit.size:
            acc += float(v) * float(q_val[hit[0]])
    return acc


def hybrid_rerank_frontend(cand_idx, cand_cos, indptr, indices, data, q_idx, q_val,
                           alpha: float, top_n: int = 10, space: str = "l2", q_sqnorm=None, d_sqnorm=None):
    """The `/search` rerank — reference `frontend/main.py:158-198` for ONE query.

    cand_idx/cand_cos: the dense top-50 (document index, cosine) in dense-rank order, i.e.
    what `collection.query(n_results=50)` returns (`:153-160`).  Chroma's default space is
    squared L2, so `semantic = 1 - dist = 2*cos - 1`; `space='cosine'`
    gives `cos`.  Empty query TF-IDF row -> all-zero TF-IDF scores (`:169-175`).  The sort
    is Python's stable `list.sort(reverse=True)` (`:197`) — ties keep dense order.
    Returns (order into the candidate list, final, semantic, tfidf), each length top_n."""
    cand_cos = np.asarray(cand_cos, dtype=np.float64)
    if space != "l2":
        sem = cand_cos
    elif q_sqnorm is None and d_sqnorm is None:
        sem = 2.0 * cand_cos - 1.0
    else:
        # hnswlib's squared-L2 distance for vectors that are not unit length (a token-less query is the zero
        # vector, `query_inferencer.py:65-69`): dist = |q|^2 + |d|^2 - 2 q.d, `semantic = 1 - dist` (`:162`)
        qn = 1.0 if q_sqnorm is None else float(q_sqnorm)
        dn = np.ones_like(cand_cos) if d_sqnorm is None else np.asarray(d_sqnorm, np.float64)
        sem = 1.0 - ((qn + dn) - 2.0 * cand_cos)
    if len(q_idx) > 0:
        tf = np.array([csr_row_dot(indptr, indices, data, int(r), np.asarray(q_idx), np.asarray(q_val))
                       for r in cand_idx], dtype=np.float64)
        tf = np.nan_to_num(tf)
    else:
        tf = np.zeros(len(cand_idx), dtype=np.float64)
    final = alpha * sem + (1.0 - alpha) * tf
    order = sorted(range(len(final)), key=lambda i: final[i], reverse=True)[:top_n]
    order = np.asarray(order, dtype=np.int64)
    return order, final[order], sem[order], tf[order]


def hybrid_search_simple(dense_cos_all, tfidf_cos_all, alpha: float, top_k: int = 10):
    """`SimpleHybridRetriever.search` — reference `backend/simple_hybrid.py:57-60`:
    corpus-wide blend then `np.argsort(combined)[::-1][:top_k]`."""
    combined = alpha * np.asarray(dense_cos_all, np.float64) + (1 - alpha) * np.asarray(tfidf_cos_all, np.float64)
    top = np.argsort(combined)[::-1][:top_k]
    return top, combined[top]


def frontend_search(query_emb, doc_emb, documents, indptr, indices, data, q_idx, q_val, alpha: float,
                    n_candidates: int = 50, n_results: int = 10, space: str = "l2"):
    """The whole `/search` handler — reference `frontend/main.py:102-210` — for one query, with the
    exact cosine top-50 standing in for `collection.query` (`:153-156`; Chroma is absent here).
    Pinned by `tests/golden/frontend_search.npz`: the responses of the unmodified handler over a
    stand-in store (`oracle/make_golden_frontend.py`).
    Returns the list that becomes `results` (without the rank / id decoration of `:206-209`)."""
    q_idx, q_val = np.asarray(q_idx), np.asarray(q_val, dtype=np.float64)
    n = len(documents)
    if alpha == 0.0:                                   # keyword branch, `:119-147`
        sims = np.array([csr_row_dot(indptr, indices, data, r, q_idx, q_val) for r in range(n)], dtype=np.float64)
        if n > n_results:
            top = np.argpartition(sims, -n_results)[-n_results:]
            top = top[np.argsort(sims[top])[::-1]]
        else:
            top = np.argsort(sims)[::-1]
        return [{"doc": documents[i], "score": float(sims[i]), "dense_score": 0.0, "tfidf_score": float(sims[i])}
                for i in top if sims[i] > 1e-5]
    kc = min(n_candidates, n)
    s, i = cosine_topk(np.asarray(query_emb, np.float32)[None, :], doc_emb, kc, dtype=np.float64)
    # a token-less query is the zero vector (`query_inferencer.py:65-69`): the store's squared-L2 distance is then
    # |d|^2, not 2 - 2cos; pinned by tests/golden/frontend_search.npz (the unmodified handler on the query "")
    q_sq = float((np.asarray(query_emb, np.float64) ** 2).sum())
    order, fin, sem, tf = hybrid_rerank_frontend(i[0], s[0], indptr, indices, data, q_idx, q_val, alpha,
                                                 top_n=min(n_results, kc), space=space,
                                                 q_sqnorm=None if abs(q_sq - 1.0) < 1e-4 else q_sq)
    return [{"doc": documents[int(i[0][o])], "score": float(f), "dense_score": float(se), "tfidf_score": float(t)}
            for o, f, se, t in zip(order, fin, sem, tf)]


def batch_eval_metrics(query_embs, doc_embs, top_k=(1, 5, 10)) -> dict:
    """`BatchEvaluator.evaluate` metrics — reference `backend/evaluators.py:49-77`: document i is the
    positive of query i; rank = position in a descending sort of row i (ties: lower index first)."""
    sim = np.asarray(query_embs, np.float64) @ np.asarray(doc_embs, np.float64).T
    n = sim.shape[0]
    ranks = np.empty(n, dtype=np.int64)
    for i in range(n):
        order = np.lexsort((np.arange(sim.shape[1]), -sim[i]))
        ranks[i] = int(np.nonzero(order == i)[0][0]) + 1
    out = {f"Recall@{k}": float((ranks <= k).mean()) for k in top_k}
    out["MRR"] = float((1.0 / ranks).mean())
    return out, ranks


def corpus_eval_metrics(query_embs, doc_embs, queries, query_to_positives, documents, top_k=(1, 5, 10)) -> dict:
    """`CorpusEvaluator` metrics — reference `backend/evaluators.py:134-209` — from embeddings:
    Recall@k = found positives / positives present in the candidate pool, Hit@k = any positive found."""
    doc_set = set(documents)
    acc = {f"Recall@{k}": [] for k in top_k}
    acc.update({f"Hit@{k}": [] for k in top_k})
    _, top = cosine_topk(np.asarray(query_embs, np.float32), np.asarray(doc_embs, np.float32), max(top_k), dtype=np.float64)
    for qi, q in enumerate(queries):
        known = query_to_positives[q]
        avail = [d for d in known if d in doc_set]
        if not avail:
            continue
        for k in top_k:
            found = len([documents[j] for j in top[qi, :k] if documents[j] in known])
            acc[f"Recall@{k}"].append(found / len(avail))
            acc[f"Hit@{k}"].append(1 if found > 0 else 0)
    return {m: (float(np.mean(v)) if v else 0.0) for m, v in acc.items()}


# --------------------------------------------------------------------------- optimiser
def clip_grad_norm(grads: "list[np.ndarray]", max_norm: float = 1.0):
    """`torch.nn.utils.clip_grad_norm_(params, max_norm)` — reference `backend/main.py:257`:
    total = ||all grads||_2 ; coef = min(1, max_norm / (total + 1e-6))."""
    total = float(np.sqrt(sum(float((g.astype(np.float64) ** 2).sum()) for g in grads)))
    coef = min(1.0, max_norm / (total + 1e-6))
    return [g * coef for g in grads], total


def adam_step(p, g, m, v, step: int, lr: float, b1=0.9, b2=0.999, eps=1e-8):
    """`torch.optim.Adam` defaults (no weight decay, no amsgrad) — reference
    `backend/main.py:222,259`.  `step` is 1-based."""
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
    p = p - (lr / bc1) * m / (np.sqrt(v) / np.sqrt(bc2) + eps)
    return p, m, v
