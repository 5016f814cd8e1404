"""Vocabulary lookup feeding int64 ids to the towers (host-side string work).

Behavioural twin of the reference `backend/tokenizer.py:6-71`: lower-case, split with the
regex `\\w+|[.,!?;]`, map out-of-vocabulary words to `<UNK>`, which is appended to the
vocabulary at index len(vocab) when the pickle does not already contain it.
"""
from __future__ import annotations

import pickle
import re
from typing import Dict, Iterable, List

_TOKEN_RE = re.compile(r"\w+|[.,!?;]")
UNK = "<UNK>"


class PretrainedTokenizer:
    def __init__(self, word_to_idx_path: str):
        with open(word_to_idx_path, "rb") as fh:          # FileNotFoundError propagates like the reference
            self.word2idx: Dict[str, int] = pickle.load(fh)
        self.unk_token = UNK
        if UNK not in self.word2idx:                      # tokenizer.py:20-24
            self.word2idx[UNK] = len(self.word2idx)
        self.unk_token_id = self.word2idx[UNK]
        self.idx2word = {i: w for w, i in self.word2idx.items()}

    def encode(self, sentence: str) -> List[int]:
        get, unk = self.word2idx.get, self.unk_token_id
        return [get(tok, unk) for tok in _TOKEN_RE.findall(str(sentence).lower())]

    def encode_batch(self, sentences: Iterable[str]) -> List[List[int]]:
        return [self.encode(s) for s in sentences]

    def decode(self, token_ids: Iterable[int]) -> str:
        return " ".join(self.idx2word.get(i, UNK) for i in token_ids)

    def vocab_size(self) -> int:
        return len(self.word2idx)

    def get_word_index(self, word: str) -> int:
        return self.word2idx.get(word, -1)

    def get_index_word(self, index: int) -> str:
        return self.idx2word.get(index, UNK)

    def contains_word(self, word: str) -> bool:
        return word in self.word2idx
