"""The reference path through the same torch call sites, functional over a state dict
(test oracle + reported CPU baseline; see oracle/__init__.py).

Differences from the reference source are structural only (functions over a dict of
tensors instead of nn.Module classes, optional dropout-mask replay); every numerical
call is the one the reference makes.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F
from torch.nn.utils.rnn import pack_padded_sequence, pad_packed_sequence


def to_torch_state(sd_numpy: dict, requires_grad: bool = False) -> Dict[str, torch.Tensor]:
    """numpy state dict -> float32 torch tensors.  The two towers' embedding tables are
    separate parameters in the reference (`backend/model.py:96-97`), so each key gets its
    own tensor.  With `requires_grad`, everything but `*.embedding.weight` is a leaf
    (pretrained tables are frozen, `model.py:25-27`)."""
    out = {}
    for k, v in sd_numpy.items():
        t = torch.tensor(v, dtype=torch.float32)
        if requires_grad and not k.endswith("embedding.weight"):
            t.requires_grad_(True)
        out[k] = t
    return out


def _gru_layer(packed, sd, prefix, layer: int, bi: bool):
    """One (bi)directional layer over a PackedSequence via `torch._VF.gru`, the call
    `nn.GRU.forward` makes for packed input (torch `nn/modules/rnn.py`, reached from
    reference `backend/model.py:59-62`)."""
    H = sd[f"{prefix}.rnn.weight_hh_l{layer}"].shape[1]
    flat = []
    for sfx in [""] + (["_reverse"] if bi else []):
        flat += [sd[f"{prefix}.rnn.weight_ih_l{layer}{sfx}"], sd[f"{prefix}.rnn.weight_hh_l{layer}{sfx}"],
                 sd[f"{prefix}.rnn.bias_ih_l{layer}{sfx}"], sd[f"{prefix}.rnn.bias_hh_l{layer}{sfx}"]]
    nb = int(packed.batch_sizes[0])
    hx = torch.zeros((2 if bi else 1), nb, H, dtype=packed.data.dtype, device=packed.data.device)
    # `train` only selects dropout (0.0 here) on CPU; cuDNN additionally refuses a backward pass after train=False
    train = torch.is_grad_enabled() and any(t.requires_grad for t in flat)
    out, hn = torch._VF.gru(packed.data, packed.batch_sizes, hx, flat, True, 1, 0.0, train, bi)
    return out, hn


def encoder_forward(sd: Dict[str, torch.Tensor], prefix: str, x: torch.Tensor, cfg: dict,
                    dropout_masks: Optional[Sequence[torch.Tensor]] = None) -> torch.Tensor:
    """`RNNEncoder.forward` — reference `backend/model.py:48-75` (eval mode, or train mode
    with the inter-layer dropout masks supplied: masks[l] is [B, T, dirs*H], already scaled
    by 1/(1-p), applied to layer l's outputs as nn.GRU does)."""
    L = cfg.get("NUM_LAYERS", 1)
    bi = bool(cfg.get("BIDIRECTIONAL", False))
    emb = F.embedding(x, sd[f"{prefix}.embedding.weight"], padding_idx=0)     # model.py:49
    lengths = (x != 0).sum(dim=1).cpu()                                      # model.py:52
    packed = pack_padded_sequence(emb, lengths, batch_first=True, enforce_sorted=False)  # :55-57
    data = packed.data
    hn = None
    for layer in range(L):
        layer_in = torch.nn.utils.rnn.PackedSequence(data, packed.batch_sizes,
                                                     packed.sorted_indices, packed.unsorted_indices)
        data, hn = _gru_layer(layer_in, sd, prefix, layer, bi)
        if dropout_masks is not None and layer < L - 1:
            po = torch.nn.utils.rnn.PackedSequence(data, packed.batch_sizes,
                                                   packed.sorted_indices, packed.unsorted_indices)
            padded, _ = pad_packed_sequence(po, batch_first=True, total_length=x.shape[1])
            padded = padded * dropout_masks[layer]
            data = pack_padded_sequence(padded, lengths, batch_first=True, enforce_sorted=False).data
    hn = hn.index_select(1, packed.unsorted_indices)                          # permute_hidden
    if bi:                                                                    # model.py:65-69
        hidden = torch.cat([hn[-2], hn[-1]], dim=1)
        hidden = F.linear(hidden, sd[f"{prefix}.projection.weight"], sd[f"{prefix}.projection.bias"])
    else:                                                                     # model.py:70-71
        hidden = hn[-1]
    if cfg.get("NORMALIZE_OUTPUT", True):                                     # model.py:73-74
        return F.normalize(hidden, p=2, dim=1)
    return hidden


def triplet_loss_cosine(q, p, n, margin: float = 0.2) -> torch.Tensor:
    """reference `backend/model.py:109-114`."""
    return torch.clamp(F.cosine_similarity(q, n) - F.cosine_similarity(q, p) + margin, min=0.0).mean()


def train_step(sd: Dict[str, torch.Tensor], opt_state: dict, cfg: dict,
               q_ids, p_ids, n_ids, masks: Optional[dict] = None,
               max_norm: Optional[float] = 1.0) -> Tuple[float, float]:
    """One live training step — reference `backend/main.py:244-259`: three encodes, cosine
    triplet loss (margin = cfg MARGIN, default 0.2), backward, `clip_grad_norm_(1.0)`, Adam.
    `max_norm=None` reproduces the un-clipped `trainer.py:101-110` variant.
    `sd` leaves are updated in place; opt_state holds {'opt': torch.optim.Adam}.
    Returns (loss, total_grad_norm_before_clipping)."""
    params = [t for k, t in sd.items() if t.requires_grad]
    if "opt" not in opt_state:
        opt_state["opt"] = torch.optim.Adam(params, lr=cfg.get("LR", 1e-4))     # main.py:222
    opt = opt_state["opt"]
    opt.zero_grad()
    m = masks or {}
    qe = encoder_forward(sd, "query_encoder", q_ids, cfg, m.get("q"))
    pe = encoder_forward(sd, "doc_encoder", p_ids, cfg, m.get("p"))
    ne = encoder_forward(sd, "doc_encoder", n_ids, cfg, m.get("n"))
    loss = triplet_loss_cosine(qe, pe, ne, margin=cfg.get("MARGIN", 0.2))         # main.py:253
    loss.backward()
    for t in params:                      # unused params (none here) would have grad None
        if t.grad is None:
            t.grad = torch.zeros_like(t)
    if max_norm is not None:
        total = float(torch.nn.utils.clip_grad_norm_(params, max_norm=max_norm))  # main.py:257
    else:
        total = float(torch.sqrt(sum((t.grad.double() ** 2).sum() for t in params)))
    opt.step()                                                                    # main.py:259
    return float(loss.detach()), total


def cosine_topk(Q: torch.Tensor, D: torch.Tensor, k: int):
    """`torch.matmul(q, D.t())` + `torch.topk` — reference `backend/evaluators.py:185-186`."""
    sim = torch.matmul(Q, D.t())
    return torch.topk(sim, k=min(k, D.shape[0]), dim=1)


def bulk_encode_documents(sd, cfg, rows: "list[list[int]]", batch_size: int = 64) -> torch.Tensor:
    """The artefact writer's encode loop — reference `backend/main.py:125-133`: batches of
    BATCH_SIZE, `pad_sequence(..., padding_value=0)`, `encode_document`, stack."""
    from torch.nn.utils.rnn import pad_sequence
    outs = []
    with torch.no_grad():
        for i in range(0, len(rows), batch_size):
            toks = [torch.tensor(r, dtype=torch.long) for r in rows[i:i + batch_size]]
            padded = pad_sequence(toks, batch_first=True, padding_value=0)
            outs.append(encoder_forward(sd, "doc_encoder", padded, cfg))
    return torch.cat(outs)
