#!/usr/bin/env python3
"""BASELINE config 1 fixture: ONE EPOCH of the reference training loop on CPU.

    python oracle/make_golden_config1.py      # needs /root/reference; ~10 min of CPU (157 steps at config.json dims)

Test infrastructure (see oracle/__init__.py).  Drives the UNMODIFIED reference classes —
`backend/model.py` (TwoTowerModel, triplet_loss_cosine), `backend/tokenizer.py`
(PretrainedTokenizer) and `backend/main.py` (TripletDataset, collate_fn; imported with an empty stub for
the absent `fastparquet`) — through the statements of the live loop `backend/main.py:244-259`
(zero_grad, three encodes, loss, backward, clip_grad_norm_(1.0), Adam(lr=LR) step) over 10,000
synthetic MS-MARCO-shaped triplets: backend/config.json dims (GRU, 2 layers, bidirectional, H=256,
E=200, batch 64, LR 5e-5, margin 0.5) with DROPOUT=0 (train-mode dropout is stochastic),
a 20,001-word synthetic vocabulary and a frozen N(0, 0.4^2) table.  `main.py`'s `DataLoader(shuffle=True)`
is replaced by an explicit seeded permutation so the GPU test can replay the same batches.
Writes tests/golden/config1_epoch.npz: the per-step loss and pre-clip gradient norm curves and the
final trainable parameters; everything else is re-created from seeds (`config1_inputs`)."""
from __future__ import annotations

import json
import os
import pickle
import sys
import tempfile
import time
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
REF = Path(os.environ.get("TTR_REFERENCE", "/root/reference"))
sys.path.insert(0, str(ROOT))

from twotowermlretrieval_b200 import synth  # noqa: E402

N_WORDS, N_TRIPLETS, BATCH = 20000, 10000, 64


def config1_inputs():
    """(cfg, word list, triplet strings, batch permutation, state dict) — all from fixed seeds; used by this
    generator AND by tests/test_gpu_train.py so both sides see identical inputs."""
    words = [f"w{i}" for i in range(N_WORDS)]
    cfg = synth.default_config(vocab_size=N_WORDS + 1, embed_dim=200)        # + '<UNK>' (tokenizer.py:20-24)
    cfg["DROPOUT"] = 0.0
    q_ids, q_len = synth.make_tokens(N_TRIPLETS, "query", N_WORDS, seed=101)
    p_ids, p_len = synth.make_tokens(N_TRIPLETS, "passage", N_WORDS, seed=102)
    n_ids, n_len = synth.make_tokens(N_TRIPLETS, "passage", N_WORDS, seed=103)

    def text(ids, ln):
        return [" ".join(words[t] for t in ids[i, :ln[i]]) for i in range(len(ln))]

    triplets = list(zip(text(q_ids, q_len), text(p_ids, p_len), text(n_ids, n_len)))
    perm = np.random.default_rng(104).permutation(N_TRIPLETS)
    sd = synth.make_state_dict(cfg, seed=105, table_seed=106)
    return cfg, words, triplets, perm, sd


SUBSAMPLE = 97      # every 97th element of the large tensors is kept (the full set is 16 MB)


def reduce_weights(final: dict, initial: dict) -> dict:
    """Final trainable parameters in a commit-sized form: a strided subsample `w::name` (stride SUBSAMPLE over the
    flattened tensor; tensors below 4,096 elements in full) and `d::name` = the same subsample of (final - initial),
    the 157-step parameter movement, which is what a parity test can meaningfully compare at LR 5e-5."""
    out = {}
    for k, v in final.items():
        stride = 1 if v.size < 4096 else SUBSAMPLE
        out[f"w::{k}"] = v.reshape(-1)[::stride].astype(np.float32)
        out[f"d::{k}"] = (v.reshape(-1).astype(np.float64) - initial[k].reshape(-1).astype(np.float64))[::stride].astype(np.float32)
    return out


def main():
    sys.path.insert(0, str(REF / "backend"))
    sys.modules.setdefault("fastparquet", types.ModuleType("fastparquet"))
    os.environ.setdefault("WANDB_MODE", "disabled")
    import main as ref_main                      # /root/reference/backend/main.py, unmodified
    import model as ref_model
    import tokenizer as ref_tok
    cfg, words, triplets, perm, sd = config1_inputs()
    with tempfile.TemporaryDirectory() as td:
        p = Path(td) / "word_to_idx.pkl"
        with open(p, "wb") as f:
            pickle.dump({w: i for i, w in enumerate(words)}, f)
        tok = ref_tok.PretrainedTokenizer(str(p))
    assert tok.vocab_size() == cfg["VOCAB_SIZE"]
    model = ref_model.TwoTowerModel(cfg, sd["query_encoder.embedding.weight"])
    model.load_state_dict({k: torch.tensor(v) for k, v in sd.items()})
    device = torch.device("cpu")
    model.to(device)
    ds = ref_main.TripletDataset(triplets, tok)
    optimizer = torch.optim.Adam(model.parameters(), lr=cfg.get("LR", 1e-4))          # main.py:222
    losses, norms = [], []
    n_steps = (N_TRIPLETS + BATCH - 1) // BATCH
    model.train()
    t0 = time.time()
    for i in range(n_steps):                                                         # main.py:244-259
        batch = [ds[int(j)] for j in perm[i * BATCH:(i + 1) * BATCH]]
        queries, pos_docs, neg_docs = ref_main.collate_fn(batch)
        queries, pos_docs, neg_docs = queries.to(device), pos_docs.to(device), neg_docs.to(device)
        optimizer.zero_grad()
        query_emb = model.encode_query(queries)
        pos_emb = model.encode_document(pos_docs)
        neg_emb = model.encode_document(neg_docs)
        loss = ref_model.triplet_loss_cosine((query_emb, pos_emb, neg_emb), margin=cfg.get("MARGIN", 0.2))
        loss.backward()
        gn = torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=1.0)
        optimizer.step()
        losses.append(loss.item())
        norms.append(float(gn))
        if (i + 1) % 10 == 0:
            print(f"step {i + 1}/{n_steps} loss {losses[-1]:.5f} |g| {norms[-1]:.4f}  ({time.time() - t0:.0f} s)", flush=True)
    final = reduce_weights({k: v.detach().numpy() for k, v in model.state_dict().items() if "embedding" not in k}, sd)
    out = ROOT / "tests" / "golden" / "config1_epoch.npz"
    np.savez_compressed(out, cfg=json.dumps(cfg), losses=np.array(losses, np.float64), grad_norms=np.array(norms, np.float64),
                        n_steps=n_steps, torch_version=torch.__version__, **final)
    print("config 1 epoch ok:", out, "avg loss", float(np.mean(losses)))


if __name__ == "__main__":
    main()
