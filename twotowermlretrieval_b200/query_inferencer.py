"""Drop-in `QueryInferencer` — reference `backend/query_inferencer.py:20-82` — on the CUDA towers.

Loads `config.json`, `word_to_idx.pkl` and `model.pth` from an artefact directory written by
the reference's `save_inference_artifacts` (`backend/main.py:92-153`) or by this package and
serves `get_query_embedding(str) -> np.float32[H]`.  Unlike the reference it does not read
`frontend/config.json` at import time, and it needs a CUDA device.
"""
from __future__ import annotations

import json
from pathlib import Path
from typing import List, Optional, Sequence

import numpy as np
import torch

from .model import TwoTowerModel
from .tokenizer import PretrainedTokenizer


class QueryInferencer:
    def __init__(self, artifacts_path: str, device: Optional[torch.device] = None):
        self.artifacts_path = Path(artifacts_path)
        self.device = torch.device(device) if device is not None else self._get_best_device()
        with open(self.artifacts_path / "config.json", "r") as fh:
            self.config = json.load(fh)
        self.tokenizer = PretrainedTokenizer(str(self.artifacts_path / "word_to_idx.pkl"))
        self.config["VOCAB_SIZE"] = self.tokenizer.vocab_size()            # query_inferencer.py:44
        self.config.setdefault("EMBED_DIM", 200)                           # query_inferencer.py:47-48
        self.model = TwoTowerModel(self.config, pretrained_embeddings=None)
        state = torch.load(self.artifacts_path / "model.pth", map_location="cpu")
        self.model.load_state_dict(state)
        self.model.to(self.device)
        self.model.device = self.device
        self.model.eval()

    def get_query_embedding(self, query: str) -> np.ndarray:
        """fp32 [H]; the zero vector when the query has no tokens (query_inferencer.py:65-69)."""
        ids = self.tokenizer.encode(query)
        if not ids:
            return np.zeros(self.config.get("HIDDEN_DIM", 128), dtype=np.float32)
        with torch.no_grad():
            x = torch.tensor(ids, dtype=torch.long).unsqueeze(0).to(self.device)
            return self.model.encode_query(x).cpu().numpy().squeeze(0)

    def encode_queries(self, queries: Sequence[str]) -> torch.Tensor:
        """Additive batched variant: fp32 [len(queries), H] on the device; token-less queries give
        zero rows.  One padded batch, one launch sequence."""
        H = self.config.get("HIDDEN_DIM", 128)
        rows = self.tokenizer.encode_batch(queries)
        out = torch.zeros(len(rows), H, dtype=torch.float32, device=self.device)
        keep = [i for i, r in enumerate(rows) if r]
        if keep:
            from .encode import encode_rows
            emb = encode_rows(self.model.query_encoder, [rows[i] for i in keep], self.device)
            out.index_copy_(0, torch.tensor(keep, device=self.device), emb)
        return out

    def _get_best_device(self) -> torch.device:
        if not torch.cuda.is_available():
            raise RuntimeError("twotowermlretrieval_b200 needs a CUDA (sm_100) device; there is no CPU path")
        return torch.device("cuda")
