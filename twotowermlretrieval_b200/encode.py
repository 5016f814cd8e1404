"""Bulk document/query encoding: the artefact writer's loop (reference
`backend/main.py:125-133`, `backend/evaluators.py:162-175`, `:242-250`) re-shaped for a GPU.

The reference encodes 64 strings at a time and copies every batch back to the host.  Here
rows are sorted by length on the host (the host tokenised them, so lengths are free),
packed into large padded batches of similar length, staged through TWO persistent pinned
buffers with asynchronous H2D copies on a side stream (the host fills batch i+1 while the
copy engine moves batch i and the SMs encode batch i-1), encoded with the zero-length check
disabled (no device->host sync per batch), and written straight into the resident output
matrix — e.g. a rank's shard of the search index — in caller order.

`streams` > 1 alternates consecutive batches between several compute streams.  Measured on a B200 (r2,
`tools/encode_rows_profile.py`, 400 k passages) that is SLOWER — 0.98-1.18 M passages/s on one stream, 0.41-0.76 M
on two, 0.41 M on three: the recurrence needs 8 free SMs of one GPC per cluster, and a second batch's one-CTA-per-SM
GEMM / gather CTAs fragment exactly those — so the default stays one stream; the switch is kept for the A/B.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import _lib, towers

Ragged = Tuple[np.ndarray, np.ndarray]      # (flat int64 token ids, int64 lengths)


ROW_QUANTUM = 3840     # rows of two full waves of the tcgen05 recurrence (15 co-resident 8-CTA clusters x 256 rows / 2 directions)


def plan_batches(lengths: np.ndarray, max_tokens: int, max_rows: int, row_quantum: int = ROW_QUANTUM):
    """Greedy batches over rows sorted by length (descending): each batch holds at most
    `max_rows` rows and `max_tokens` padded tokens; row counts above `row_quantum` are rounded down to a
    multiple of it so the recurrence kernel runs whole waves of clusters.  Returns (order, [(lo, hi), ...])."""
    mx = int(lengths.max()) if len(lengths) else 0
    if 0 <= int(lengths.min() if len(lengths) else 0) and mx < 65536:
        # 16-bit keys: numpy's stable sort is a radix sort there (19 ms instead of 127 ms per 1.1 M rows)
        order = np.argsort((mx - lengths).astype(np.uint16), kind="stable")
    else:
        order = np.argsort(-lengths, kind="stable")
    sl = lengths[order]
    bounds, lo, n = [], 0, len(sl)
    while lo < n:
        T = int(sl[lo])
        rows = max(1, min(max_rows, max_tokens // max(T, 1)))
        if row_quantum > 0 and rows > row_quantum:
            rows -= rows % row_quantum
        hi = min(n, lo + rows)
        bounds.append((lo, hi))
        lo = hi
    return order, bounds


def to_ragged(rows: Union[Ragged, Sequence[Sequence[int]]]) -> Ragged:
    """List of token-id lists (what `PretrainedTokenizer.encode` returns per string) -> one flat
    int64 array + lengths; a (flat, lengths) pair passes through."""
    if isinstance(rows, tuple) and len(rows) == 2 and isinstance(rows[0], np.ndarray):
        return np.ascontiguousarray(rows[0], dtype=np.int64), np.asarray(rows[1], dtype=np.int64)
    n = len(rows)
    lengths = np.fromiter((len(r) for r in rows), dtype=np.int64, count=n)
    flat = np.empty(int(lengths.sum()), dtype=np.int64)
    o = 0
    for r in rows:
        m = len(r)
        flat[o:o + m] = r
        o += m
    return flat, lengths


class _Staging:
    """One pinned (ids, out-row index) buffer pair and the event of the last H2D copy out of it."""

    def __init__(self, max_tokens: int, max_rows: int):
        self.ids = torch.empty(max_tokens, dtype=torch.int64).pin_memory()
        self.idx = torch.empty(max_rows, dtype=torch.int64).pin_memory()
        self.ids_np, self.idx_np = self.ids.numpy(), self.idx.numpy()
        self.copied: Optional[torch.cuda.Event] = None


class _Lanes:
    """`n` compute streams that take turns; `begin()` makes them wait for the caller's stream (weights, earlier
    work), `end()` makes the caller's stream wait for all of them.  n = 1 is the caller's stream itself."""

    def __init__(self, encoder, device, n: int):
        self.dev = torch.device(device)
        self.main = torch.cuda.current_stream(self.dev)
        self.streams = [self.main] if n <= 1 else [torch.cuda.Stream(device=self.dev) for _ in range(n)]
        # everything a lane reads must be complete on the caller's stream first: the flat parameter buffer and the
        # fp16 copies of W_ih are (re)built lazily by the first forward that needs them
        encoder._ensure_flat()
        if encoder.hidden_dim == 256 and towers._fp16_pipeline_allowed():
            for layer in range(encoder.num_layers):
                towers._w16_cached(encoder, layer, encoder.layer_weights(layer)[0])
        for s in self.streams:
            if s is not self.main:
                s.wait_stream(self.main)
        self.i = 0

    def next(self) -> "torch.cuda.Stream":
        s = self.streams[self.i % len(self.streams)]
        self.i += 1
        return s

    def end(self):
        for s in self.streams:
            if s is not self.main:
                self.main.wait_stream(s)


def encode_padded_batches(encoder, batches: Sequence[torch.Tensor], streams: int = 1) -> "list[torch.Tensor]":
    """Encode device-resident padded id batches [R_i, T_i] (inference, zero-length check off), optionally with
    consecutive batches on alternating compute streams; returns the embeddings in order.  Results are ordered after the
    caller's current stream like any other call."""
    if not batches:
        return []
    was_strict, was_training = encoder.strict_lengths, encoder.training
    encoder.strict_lengths = False
    encoder.eval()
    lanes = _Lanes(encoder, batches[0].device, streams)
    outs = []
    try:
        with torch.no_grad():
            for b in batches:
                s = lanes.next()
                with torch.cuda.stream(s):
                    e = encoder(b)
                if s is not lanes.main:
                    e.record_stream(lanes.main)
                outs.append(e)
    finally:
        lanes.end()
        encoder.strict_lengths = was_strict
        encoder.train(was_training)
    return outs


def encode_rows(encoder, rows: Union[Ragged, Sequence[Sequence[int]]], device, out: Optional[torch.Tensor] = None,
                out_offset: int = 0, max_tokens: int = 1048576, max_rows: int = 30720, streams: int = 1) -> torch.Tensor:
    """Encode tokenised rows with `encoder` (an RNNEncoder) -> fp32 [n, H] on `device`
    (rows `out[out_offset : out_offset + n]` if `out` is given).  Batches hold up to `max_tokens` padded tokens /
    `max_rows` rows (r2, 600 k passages end to end: 1.59 M passages/s at 524,288 / 15,360, 1.76 M at 1,048,576 /
    30,720 — eight waves of the recurrence per launch instead of four; ~5 GB of fp16 gi / y / X in flight).  `rows` is a list of id lists or a
    (flat ids, lengths) pair.  Raises RuntimeError for empty rows like the reference's
    pack_padded_sequence (quirk #2)."""
    flat, lengths = to_ragged(rows)
    n = len(lengths)
    H = encoder.hidden_dim
    if out is None:
        out = torch.empty(n, H, dtype=torch.float32, device=device)
        out_offset = 0
    if n == 0:
        return out
    from .model import _ZERO_LEN_MSG
    starts = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(lengths, out=starts[1:])
    # quirk #1: the effective length is the count of non-zero ids; rows are passed through untouched, the device
    # plan recounts.  Only the all-zero / empty case is an error; the packer counts non-zero ids while it copies
    # (no separate pass over the corpus before the first batch), so an all-zero row raises when its batch is packed.
    if (lengths <= 0).any():
        raise RuntimeError(_ZERO_LEN_MSG)
    order, bounds = plan_batches(lengths, max_tokens, max_rows)
    import ctypes
    nnz_total, zero_rows = ctypes.c_int64(0), ctypes.c_int64(0)
    max_tok = max((hi - lo) * int(lengths[order[lo]]) for lo, hi in bounds)
    max_row = max(hi - lo for lo, hi in bounds)
    stage = [_Staging(max_tok, max_row) for _ in range(2)]
    was_strict, was_training = encoder.strict_lengths, encoder.training
    encoder.strict_lengths = False
    encoder.eval()
    dev = torch.device(device)
    copy_stream = torch.cuda.Stream(device=dev)
    lanes = _Lanes(encoder, dev, streams)
    try:
        with torch.no_grad():
            for bi, (lo, hi) in enumerate(bounds):
                st = stage[bi & 1]
                if st.copied is not None:
                    st.copied.synchronize()              # the copy engine has read this buffer (two batches ago)
                idx = order[lo:hi]
                R, T = hi - lo, int(lengths[idx[0]])
                # fill the padded [R, T] view of the pinned buffer from the ragged rows (host memcpy per row, in C)
                idx = np.ascontiguousarray(idx, dtype=np.int64)
                _lib.call_nostream("ttr_pack_padded_count_i64", flat.ctypes.data, starts.ctypes.data, lengths.ctypes.data,
                                   idx.ctypes.data, R, T, st.ids.data_ptr(), ctypes.addressof(nnz_total),
                                   ctypes.addressof(zero_rows))
                if zero_rows.value:
                    raise RuntimeError(_ZERO_LEN_MSG)
                st.idx_np[:R] = idx + out_offset
                with torch.cuda.stream(copy_stream):
                    dev_ids = st.ids[:R * T].view(R, T).to(dev, non_blocking=True)
                    dev_idx = st.idx[:R].to(dev, non_blocking=True)
                    st.copied = torch.cuda.Event()
                    st.copied.record(copy_stream)
                cs = lanes.next()
                cs.wait_stream(copy_stream)
                encoder._token_bound = int(nnz_total.value)     # rows of the packed per-token matrices (<= R * T)
                with torch.cuda.stream(cs):
                    emb = encoder(dev_ids)
                    out.index_copy_(0, dev_idx, emb)            # distinct rows per batch: lanes never write the same row
                dev_ids.record_stream(cs)
                dev_idx.record_stream(cs)
    finally:
        lanes.end()
        encoder._token_bound = None
        encoder.strict_lengths = was_strict
        encoder.train(was_training)
    return out


def encode_texts(encoder, tokenizer, texts: Sequence[str], device, **kw) -> torch.Tensor:
    return encode_rows(encoder, tokenizer.encode_batch(texts), device, **kw)
