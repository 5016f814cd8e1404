"""Triplet batches for the training / evaluation loops — reference `backend/main.py:33-56`.

Same `TripletDataset` / `collate_fn` contract ((query, positive, negative) strings -> three int64
[B, T] tensors, right-padded with 0 to the batch maximum).  Tokenisation and padding happen for a whole
batch at once on the host (one numpy fill per tensor) instead of one `torch.tensor` + `pad_sequence`
per sample.  The device side of the input pipeline is `encode.encode_rows`.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np
import torch


def pad_rows(rows: Sequence[Sequence[int]]) -> torch.Tensor:
    """int64 [len(rows), max_len], zero right-padded — `pad_sequence(batch_first=True, padding_value=0)`."""
    T = max((len(r) for r in rows), default=0)
    out = np.zeros((len(rows), T), dtype=np.int64)
    for i, r in enumerate(rows):
        out[i, :len(r)] = r
    return torch.from_numpy(out)


class TripletDataset(torch.utils.data.Dataset):
    """`backend/main.py:33-48`: item i -> three 1-D int64 token tensors."""

    def __init__(self, data: List[Tuple[str, str, str]], tokenizer):
        self.data = data
        self.tokenizer = tokenizer

    def __len__(self) -> int:
        return len(self.data)

    def __getitem__(self, idx: int):
        q, p, n = self.data[idx]
        enc = self.tokenizer.encode
        return (torch.tensor(enc(q), dtype=torch.long), torch.tensor(enc(p), dtype=torch.long),
                torch.tensor(enc(n), dtype=torch.long))


def collate_fn(batch):
    """`backend/main.py:50-56`."""
    queries, pos_docs, neg_docs = zip(*batch)
    return (pad_rows([t.tolist() for t in queries]), pad_rows([t.tolist() for t in pos_docs]),
            pad_rows([t.tolist() for t in neg_docs]))


def triplet_batches(data: List[Tuple[str, str, str]], tokenizer, batch_size: int = 64):
    """Batched equivalent of DataLoader(TripletDataset, collate_fn, shuffle=False): tokenises each string once."""
    for lo in range(0, len(data), batch_size):
        chunk = data[lo:lo + batch_size]
        yield (pad_rows([tokenizer.encode(q) for q, _, _ in chunk]), pad_rows([tokenizer.encode(p) for _, p, _ in chunk]),
               pad_rows([tokenizer.encode(n) for _, _, n in chunk]))
