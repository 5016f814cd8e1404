"""Deterministic synthetic inputs for parity tests and benchmarks (numpy only).

Everything here is seeded with numpy's PCG64 streams so the same arrays can be
re-created on the GPU box, in `oracle/make_golden.py` (which feeds them to the
unmodified reference) and in `bench.py`.  Shapes and distributions follow
SURVEY.md §8(d): GRU/Linear parameters U(-1/sqrt(fan), 1/sqrt(fan)) like torch's
default init (reference `backend/model.py:31-46`), a GloVe-like N(0, 0.4^2)
embedding table, Zipf(1.07) token ids over [1, V), MS-MARCO-shaped lengths.
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "tower_param_shapes", "make_state_dict", "make_tokens", "make_ragged_tokens", "make_lengths",
    "pad_rows", "make_unit_rows", "make_tfidf_csr", "default_config",
]


def default_config(vocab_size: int = 400005, embed_dim: int = 200) -> dict:
    """`backend/config.json:13-24` plus the two runtime keys `main.py:184-185` adds."""
    return {
        "VOCAB_SIZE": vocab_size, "EMBED_DIM": embed_dim, "HIDDEN_DIM": 256,
        "RNN_TYPE": "GRU", "NUM_LAYERS": 2, "BIDIRECTIONAL": True, "DROPOUT": 0.2,
        "BATCH_SIZE": 64, "EPOCHS": 1, "LR": 5e-5, "MARGIN": 0.5,
        "NORMALIZE_OUTPUT": True,
    }


def tower_param_shapes(cfg: dict) -> "dict[str, tuple]":
    """Parameter names/shapes of one tower, in torch registration order
    (embedding, GRU flat weights per layer/direction, projection) — the
    state_dict contract of `backend/model.py:24-46`."""
    V, E, H = cfg["VOCAB_SIZE"], cfg["EMBED_DIM"], cfg["HIDDEN_DIM"]
    L = cfg.get("NUM_LAYERS", 1)
    bi = bool(cfg.get("BIDIRECTIONAL", False))
    shapes = {"embedding.weight": (V, E)}
    for layer in range(L):
        in_dim = E if layer == 0 else H * (2 if bi else 1)
        for sfx in ([""] + (["_reverse"] if bi else [])):
            shapes[f"rnn.weight_ih_l{layer}{sfx}"] = (3 * H, in_dim)
            shapes[f"rnn.weight_hh_l{layer}{sfx}"] = (3 * H, H)
            shapes[f"rnn.bias_ih_l{layer}{sfx}"] = (3 * H,)
            shapes[f"rnn.bias_hh_l{layer}{sfx}"] = (3 * H,)
    if bi:
        shapes["projection.weight"] = (H, 2 * H)
        shapes["projection.bias"] = (H,)
    return shapes


def make_state_dict(cfg: dict, seed: int = 0, table_seed: int = 1,
                    table_std: float = 0.4, zero_pad_row: bool = False) -> "dict[str, np.ndarray]":
    """Both towers' parameters as float32 numpy arrays under the reference key names
    (`query_encoder.*`, `doc_encoder.*`).  The two towers share one table (the
    reference copies the same GloVe matrix into both, `model.py:25-26,96-97`) but have
    independent GRU/projection weights."""
    H = cfg["HIDDEN_DIM"]
    rng = np.random.default_rng(seed)
    trng = np.random.default_rng(table_seed)
    table = (trng.standard_normal((cfg["VOCAB_SIZE"], cfg["EMBED_DIM"]), dtype=np.float32)
             * np.float32(table_std))
    if zero_pad_row:
        table[0] = 0.0
    sd = {}
    for tower in ("query_encoder", "doc_encoder"):
        for name, shape in tower_param_shapes(cfg).items():
            if name == "embedding.weight":
                sd[f"{tower}.{name}"] = table
                continue
            bound = 1.0 / np.sqrt(2 * H if name.startswith("projection") else H)
            sd[f"{tower}.{name}"] = rng.uniform(-bound, bound, size=shape).astype(np.float32)
    return sd


def make_lengths(n: int, kind: str, rng: np.random.Generator) -> np.ndarray:
    """Query lengths clip(round(N(6,2.5)),2,30); passage lengths
    clip(round(lognormal(4.1,0.4)),8,256) (mean ~64) — SURVEY.md §8(d)."""
    if kind == "query":
        return np.clip(np.rint(rng.normal(6.0, 2.5, size=n)), 2, 30).astype(np.int64)
    if kind == "passage":
        return np.clip(np.rint(rng.lognormal(4.1, 0.4, size=n)), 8, 256).astype(np.int64)
    raise ValueError(kind)


_ZIPF_CACHE: dict = {}


def _zipf_cdf(vocab_size: int, s: float) -> np.ndarray:
    key = (vocab_size, s)
    if key not in _ZIPF_CACHE:
        w = np.arange(1, vocab_size, dtype=np.float64) ** (-s)
        _ZIPF_CACHE[key] = np.cumsum(w) / w.sum()
    return _ZIPF_CACHE[key]


def make_tokens(n: int, kind: str, vocab_size: int, seed: int = 2,
                lengths: "np.ndarray | None" = None, pad_to: "int | None" = None):
    """(ids int64 [n, T] right-padded with 0, lengths int64 [n]).  Ids are Zipf(1.07)
    ranks in [1, V) — never 0, so id-0 quirk #1 is only exercised by dedicated tests."""
    rng = np.random.default_rng(seed)
    if lengths is None:
        lengths = make_lengths(n, kind, rng)
    lengths = np.asarray(lengths, dtype=np.int64)
    cdf = _zipf_cdf(vocab_size, 1.07)
    total = int(lengths.sum())
    flat = (np.searchsorted(cdf, rng.random(total), side="left") + 1).astype(np.int64)
    np.clip(flat, 1, vocab_size - 1, out=flat)
    return pad_rows(flat, lengths, pad_to), lengths


def make_ragged_tokens(n: int, kind: str, vocab_size: int, seed: int = 2, chunk: int = 1 << 22):
    """(flat int64 ids, lengths int64 [n]) — the same distributions as `make_tokens` without the padded matrix
    (bulk-encode inputs: 1.1 M passages are 72 M tokens, their padded [n, 256] form would be 2.3 GB)."""
    rng = np.random.default_rng(seed)
    lengths = make_lengths(n, kind, rng)
    cdf = _zipf_cdf(vocab_size, 1.07)
    total = int(lengths.sum())
    flat = np.empty(total, dtype=np.int64)
    for lo in range(0, total, chunk):
        hi = min(total, lo + chunk)
        flat[lo:hi] = np.searchsorted(cdf, rng.random(hi - lo), side="left") + 1
    np.clip(flat, 1, vocab_size - 1, out=flat)
    return flat, lengths


def pad_rows(flat: np.ndarray, lengths: np.ndarray, pad_to: "int | None" = None) -> np.ndarray:
    """Ragged -> right-padded [n, T] with 0, the layout `pad_sequence(batch_first=True,
    padding_value=0)` produces in `backend/main.py:50-56`."""
    n = len(lengths)
    T = int(pad_to if pad_to is not None else (lengths.max() if n else 0))
    out = np.zeros((n, T), dtype=np.int64)
    mask = np.arange(T)[None, :] < lengths[:, None]
    out[mask] = flat
    return out


def make_unit_rows(n: int, dim: int, seed: int) -> np.ndarray:
    """F.normalize(N(0,1)) float32 rows — synthetic document/query embeddings."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, dim), dtype=np.float32)
    x /= np.maximum(np.linalg.norm(x, axis=1, keepdims=True), 1e-12)
    return x.astype(np.float32)


def make_tfidf_csr(n_rows: int, n_features: int = 20000, mean_nnz: float = 30.0,
                   seed: int = 5, min_nnz: int = 1):
    """Synthetic L2-normalised TF-IDF CSR (indptr int64, indices int32 sorted per row,
    data float64) shaped like sklearn's `TfidfVectorizer(norm='l2')` output
    (`backend/main.py:142-143`)."""
    rng = np.random.default_rng(seed)
    nnz = np.maximum(rng.poisson(mean_nnz, size=n_rows), min_nnz).astype(np.int64)
    nnz = np.minimum(nnz, n_features)
    indptr = np.zeros(n_rows + 1, dtype=np.int64)
    np.cumsum(nnz, out=indptr[1:])
    indices = np.empty(int(indptr[-1]), dtype=np.int32)
    data = np.empty(int(indptr[-1]), dtype=np.float64)
    for r in range(n_rows):
        k = int(nnz[r])
        cols = np.sort(rng.choice(n_features, size=k, replace=False)).astype(np.int32)
        vals = rng.random(k) + 0.05
        vals /= np.linalg.norm(vals)
        indices[indptr[r]:indptr[r + 1]] = cols
        data[indptr[r]:indptr[r + 1]] = vals
    return indptr, indices, data
