"""Bulk document/query encoding: the artefact writer's loop (reference
`backend/main.py:125-133`, `backend/evaluators.py:162-175`, `:242-250`) re-shaped for a GPU.

The reference encodes 64 strings at a time and copies every batch back to the host.  Here
rows are sorted by length on the host (the host tokenised them, so lengths are free),
packed into large padded batches of similar length, staged through pinned memory with
asynchronous H2D copies, encoded with the zero-length check disabled (no device->host sync
per batch), and written straight into the resident output matrix in caller order.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib


def plan_batches(lengths: np.ndarray, max_tokens: int, max_rows: int):
    """Greedy batches over rows sorted by length (descending): each batch holds at most
    `max_rows` rows and `max_tokens` padded tokens.  Returns (order, [(lo, hi), ...])."""
    order = np.argsort(-lengths, kind="stable")
    sl = lengths[order]
    bounds, lo, n = [], 0, len(sl)
    while lo < n:
        T = int(sl[lo])
        rows = max(1, min(max_rows, max_tokens // max(T, 1)))
        hi = min(n, lo + rows)
        bounds.append((lo, hi))
        lo = hi
    return order, bounds


def encode_rows(encoder, rows: Sequence[Sequence[int]], device, out: Optional[torch.Tensor] = None,
                out_offset: int = 0, max_tokens: int = 262144, max_rows: int = 16384) -> torch.Tensor:
    """Encode tokenised rows with `encoder` (an RNNEncoder) -> fp32 [len(rows), H] on `device`
    (rows of `out[out_offset:]` if given).  Raises RuntimeError for empty rows like the
    reference's pack_padded_sequence (quirk #2)."""
    n = len(rows)
    H = encoder.hidden_dim
    if out is None:
        out = torch.empty(n, H, dtype=torch.float32, device=device)
        out_offset = 0
    if n == 0:
        return out
    lengths = np.fromiter((len(r) for r in rows), dtype=np.int64, count=n)
    # quirk #1: the effective length is the count of non-zero ids; rows are passed through untouched,
    # the device plan recounts.  Only the all-zero / empty case is an error.
    nnz = np.fromiter((sum(1 for t in r if t != 0) for r in rows), dtype=np.int64, count=n)
    if (nnz <= 0).any():
        from .model import _ZERO_LEN_MSG
        raise RuntimeError(_ZERO_LEN_MSG)
    order, bounds = plan_batches(lengths, max_tokens, max_rows)
    was_strict, was_training = encoder.strict_lengths, encoder.training
    encoder.strict_lengths = False
    encoder.eval()
    copy_stream = torch.cuda.Stream(device=device)
    try:
        with torch.no_grad():
            staged = None
            for bi, (lo, hi) in enumerate(bounds):
                idx = order[lo:hi]
                T = int(lengths[idx[0]])
                host = torch.zeros(hi - lo, T, dtype=torch.int64).pin_memory()
                hv = host.numpy()
                for r, src in enumerate(idx):
                    row = rows[src]
                    hv[r, :len(row)] = row
                with torch.cuda.stream(copy_stream):
                    dev_ids = host.to(device, non_blocking=True)
                    dev_idx = torch.from_numpy(idx.astype(np.int64) + out_offset).pin_memory().to(device, non_blocking=True)
                torch.cuda.current_stream(device).wait_stream(copy_stream)
                emb = encoder(dev_ids)
                out.index_copy_(0, dev_idx, emb)
                dev_ids.record_stream(torch.cuda.current_stream(device))
                dev_idx.record_stream(torch.cuda.current_stream(device))
    finally:
        encoder.strict_lengths = was_strict
        encoder.train(was_training)
    return out


def encode_texts(encoder, tokenizer, texts: Sequence[str], device, **kw) -> torch.Tensor:
    return encode_rows(encoder, tokenizer.encode_batch(texts), device, **kw)
