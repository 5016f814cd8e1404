"""Drop-in mirror of the reference `backend/model.py` on top of the sm_100a kernels.

Same classes, constructor arguments, state_dict keys and error behaviour as the reference
(`RNNEncoder` model.py:8-75, `TwoTowerModel` model.py:78-106, `triplet_loss_cosine`
model.py:109-114); the forward/backward math runs in libttr_b200.so.  Only `RNN_TYPE='GRU'`
is implemented (the shipped config, backend/config.json:14).  There is no CPU path: calling
`forward` on CPU tensors raises.

Parameter storage: every trainable tensor of a module tree is a view into one flat fp32
buffer (`flat_params()`), laid out so that the two directions' `weight_ih` of a layer are
adjacent — the input-projection GEMM then sees a single [dirs*3H, in] matrix, and the
data-parallel all-reduce + fused clip/Adam run over one bucket.
"""
from __future__ import annotations

import math
import os
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn as nn

from . import _lib

_ZERO_LEN_MSG = ("Length of all samples has to be greater than 0, but found an element in 'lengths' "
                 "that is <= 0")


class _GRUWeights(nn.Module):
    """Parameter container with torch.nn.GRU's names, shapes, registration order and init
    (U(-1/sqrt(H), 1/sqrt(H)) in registration order) so state_dicts and seeds are interchangeable
    with the reference's `nn.GRU` (model.py:31-37)."""

    def __init__(self, input_size: int, hidden_size: int, num_layers: int, bidirectional: bool, dropout: float):
        super().__init__()
        self.input_size, self.hidden_size, self.num_layers = input_size, hidden_size, num_layers
        self.bidirectional, self.dropout = bidirectional, float(dropout)
        dirs = 2 if bidirectional else 1
        for layer in range(num_layers):
            in_dim = input_size if layer == 0 else hidden_size * dirs
            for sfx in [""] + (["_reverse"] if bidirectional else []):
                self.register_parameter(f"weight_ih_l{layer}{sfx}", nn.Parameter(torch.empty(3 * hidden_size, in_dim)))
                self.register_parameter(f"weight_hh_l{layer}{sfx}", nn.Parameter(torch.empty(3 * hidden_size, hidden_size)))
                self.register_parameter(f"bias_ih_l{layer}{sfx}", nn.Parameter(torch.empty(3 * hidden_size)))
                self.register_parameter(f"bias_hh_l{layer}{sfx}", nn.Parameter(torch.empty(3 * hidden_size)))
        stdv = 1.0 / math.sqrt(hidden_size) if hidden_size > 0 else 0
        for p in self.parameters():
            nn.init.uniform_(p, -stdv, stdv)

    def extra_repr(self):
        return (f"{self.input_size}, {self.hidden_size}, num_layers={self.num_layers}, "
                f"bidirectional={self.bidirectional}, dropout={self.dropout}")


def _flat_order(enc: "RNNEncoder") -> List[Tuple[str, nn.Parameter]]:
    """Trainable tensors of one tower in kernel order: per layer W_ih(fwd,rev), b_ih(fwd,rev),
    W_hh(fwd,rev), b_hh(fwd,rev); then projection weight, bias.  The two directions of one
    kind are packed back to back (no padding) so they read as one [dirs*3H, ...] tensor."""
    out = []
    sfxs = [""] + (["_reverse"] if enc.bidirectional else [])
    for layer in range(enc.num_layers):
        for kind in ("weight_ih", "bias_ih", "weight_hh", "bias_hh"):
            for sfx in sfxs:
                name = f"{kind}_l{layer}{sfx}"
                out.append((f"rnn.{name}", getattr(enc.rnn, name)))
    if enc.projection is not None:
        out.append(("projection.weight", enc.projection.weight))
        out.append(("projection.bias", enc.projection.bias))
    return out


def _pack_flat(named: List[Tuple[str, nn.Parameter]], keep_grad: bool = True):
    """Move the listed parameters into one contiguous buffer and re-point `.data` (and `.grad`)
    at views of it; groups (everything but a `_reverse` twin) start 16-byte aligned.
    Returns (flat, flat_grad, slices)."""
    if not named:
        return None, None, {}
    dev = named[0][1].device
    offs, total = {}, 0
    for name, p in named:
        if not name.endswith("_reverse"):
            total = (total + 3) // 4 * 4
        offs[name] = total
        total += p.numel()
    total = (total + 3) // 4 * 4
    flat = torch.zeros(total, dtype=torch.float32, device=dev)
    flat_grad = torch.zeros(total, dtype=torch.float32, device=dev)
    slices = {}
    for name, p in named:
        o, n = offs[name], p.numel()
        flat[o:o + n].copy_(p.data.reshape(-1).to(torch.float32))
        if keep_grad and p.grad is not None:
            flat_grad[o:o + n].copy_(p.grad.reshape(-1))
        p.data = flat[o:o + n].view(p.shape)
        p.grad = flat_grad[o:o + n].view(p.shape) if p.requires_grad else None
        slices[name] = (o, n)
    return flat, flat_grad, slices


class RNNEncoder(nn.Module):
    """GloVe-embedding + (bi)GRU text encoder — reference `backend/model.py:8-75`."""

    def __init__(self, vocab_size: int, embed_dim: int, hidden_dim: int,
                 pretrained_embeddings: Optional[np.ndarray] = None, rnn_type: str = "GRU",
                 num_layers: int = 1, dropout: float = 0.0, bidirectional: bool = False,
                 normalize_output: bool = True):
        super().__init__()
        if rnn_type.upper() != "GRU":
            raise NotImplementedError(f"rnn_type={rnn_type!r}: only 'GRU' (backend/config.json:14) has sm_100a kernels")
        self.embedding = nn.Embedding(vocab_size, embed_dim, padding_idx=0)
        if pretrained_embeddings is not None:
            self.embedding.weight.data.copy_(torch.from_numpy(np.asarray(pretrained_embeddings)))
            self.embedding.weight.requires_grad = False
        self.bidirectional = bool(bidirectional)
        self.num_layers = int(num_layers)
        self.hidden_dim = int(hidden_dim)
        self.rnn = _GRUWeights(embed_dim, hidden_dim, num_layers, self.bidirectional,
                               dropout if num_layers > 1 else 0.0)
        self.rnn_type = "GRU"
        self.normalize_output = bool(normalize_output)
        self.projection = nn.Linear(hidden_dim * 2, hidden_dim) if self.bidirectional else None
        # reference behaviour: zero-length rows raise (costs one 16-byte D2H read per call, like the
        # reference's `.cpu()` on lengths, model.py:52).  Bulk paths that know their lengths switch it off.
        self.strict_lengths = True
        self.last_dropout_masks: Optional[list] = None   # packed-layout masks of the last train-mode forward
        self._flat = self._flat_grad = None
        self._slices: Dict[str, Tuple[int, int]] = {}
        self._owner = None   # TwoTowerModel that owns the flat buffer, if any
        self._prefix = ""

    # ------------------------------------------------------------------ flat storage
    def flatten_parameters(self):
        if self._owner is not None:
            self._owner.flatten_parameters()
            return
        self._flat, self._flat_grad, self._slices = _pack_flat(_flat_order(self))

    def _apply(self, fn, recurse=True):
        r = super()._apply(fn, recurse)
        if self._owner is None:
            self._flat = None      # storage moved: re-pack lazily
        return r

    def _ensure_flat(self):
        if self._owner is not None:
            self._owner._ensure_flat()
        elif self._flat is None or self._flat.device != self.rnn.weight_hh_l0.device or not self._views_ok():
            self.flatten_parameters()

    def _views_ok(self) -> bool:
        base = self._flat.data_ptr()
        for name, p in _flat_order(self):
            o, _ = self._slices.get(name, (None, None))
            if o is None or p.data_ptr() != base + 4 * o:
                return False
        return True

    def _w(self, name: str, shape) -> torch.Tensor:
        owner = self._owner if self._owner is not None else self
        o, _ = owner._slices[(self._prefix + name) if self._owner is not None else name]
        n = int(np.prod(shape))
        return owner._flat[o:o + n].view(*shape)

    def layer_weights(self, layer: int):
        """(W_ih [dirs*3H, in], b_ih [dirs*3H], W_hh [dirs, 3H, H], b_hh [dirs*3H]) views."""
        dirs, H = (2 if self.bidirectional else 1), self.hidden_dim
        in_dim = self.embedding.embedding_dim if layer == 0 else dirs * H
        return (self._w(f"rnn.weight_ih_l{layer}", (dirs * 3 * H, in_dim)),
                self._w(f"rnn.bias_ih_l{layer}", (dirs * 3 * H,)),
                self._w(f"rnn.weight_hh_l{layer}", (dirs, 3 * H, H)),
                self._w(f"rnn.bias_hh_l{layer}", (dirs * 3 * H,)))

    # ------------------------------------------------------------------ forward
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        from .towers import encoder_forward
        return encoder_forward(self, x)


class TwoTowerModel(nn.Module):
    """Two independent towers — reference `backend/model.py:78-106`."""

    def __init__(self, config: Dict, pretrained_embeddings: Optional[np.ndarray] = None):
        super().__init__()
        encoder_args = dict(
            vocab_size=config["VOCAB_SIZE"], embed_dim=config["EMBED_DIM"], hidden_dim=config["HIDDEN_DIM"],
            pretrained_embeddings=pretrained_embeddings, rnn_type=config.get("RNN_TYPE", "GRU"),
            num_layers=config.get("NUM_LAYERS", 1), dropout=config.get("DROPOUT", 0.0),
            bidirectional=config.get("BIDIRECTIONAL", False),
            normalize_output=config.get("NORMALIZE_OUTPUT", True))
        self.query_encoder = RNNEncoder(**encoder_args)
        self.doc_encoder = RNNEncoder(**encoder_args)
        self.query_encoder._owner = self.doc_encoder._owner = _OwnerRef(self)
        self.query_encoder._prefix, self.doc_encoder._prefix = "query_encoder.", "doc_encoder."
        self._flat = self._flat_grad = None
        self._slices: Dict[str, Tuple[int, int]] = {}

    def _named_flat(self):
        return ([("query_encoder." + n, p) for n, p in _flat_order(self.query_encoder)] +
                [("doc_encoder." + n, p) for n, p in _flat_order(self.doc_encoder)])

    def flatten_parameters(self):
        self._flat, self._flat_grad, self._slices = _pack_flat(self._named_flat())

    def _apply(self, fn, recurse=True):
        r = super()._apply(fn, recurse)
        self._flat = None
        return r

    def _ensure_flat(self):
        if self._flat is None or not self._views_ok():
            self.flatten_parameters()

    def _views_ok(self) -> bool:
        base = self._flat.data_ptr()
        for name, p in self._named_flat():
            o, _ = self._slices.get(name, (None, None))
            if o is None or p.data_ptr() != base + 4 * o:
                return False
        return True

    def flat_params(self) -> torch.Tensor:
        """All GRU/projection parameters of both towers as one fp32 vector (views, not copies)."""
        self._ensure_flat()
        return self._flat

    def flat_grads(self) -> torch.Tensor:
        self._ensure_flat()
        for name, p in self._named_flat():     # autograd may have replaced a .grad tensor
            o, n = self._slices[name]
            if p.requires_grad and (p.grad is None or p.grad.data_ptr() != self._flat_grad.data_ptr() + 4 * o):
                if p.grad is not None:
                    self._flat_grad[o:o + n].copy_(p.grad.reshape(-1))
                p.grad = self._flat_grad[o:o + n].view(p.shape)
        return self._flat_grad

    def encode_query(self, query: torch.Tensor) -> torch.Tensor:
        return self.query_encoder(query)

    def encode_document(self, document: torch.Tensor) -> torch.Tensor:
        return self.doc_encoder(document)

    def forward(self, query: torch.Tensor, document: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        return self.encode_query(query), self.encode_document(document)


class _OwnerRef:
    """Non-module handle so a child encoder can reach its TwoTowerModel without registering a cycle."""

    def __init__(self, owner: TwoTowerModel):
        self._o = owner

    def __getattr__(self, k):
        # copy / pickle probe dunders (`__deepcopy__`, `__setstate__`, ...) on a not-yet-initialised instance
        if k.startswith("__") or "_o" not in self.__dict__:
            raise AttributeError(k)
        return getattr(self.__dict__["_o"], k)

    def __deepcopy__(self, memo):
        # `copy.deepcopy(model)` (the usual way to snapshot a best model; the reference nn.Module supports it):
        # the handle of the copy must point at the COPY of the owner, which deepcopy has already registered in
        # `memo`; an encoder copied on its own becomes a stand-alone module
        new_owner = memo.get(id(self.__dict__["_o"]))
        return _OwnerRef(new_owner) if new_owner is not None else None


def triplet_loss_cosine(triplet: Tuple[torch.Tensor, torch.Tensor, torch.Tensor], margin: float = 0.2) -> torch.Tensor:
    """Cosine triplet loss — reference `backend/model.py:109-114` (differentiable)."""
    from .towers import TripletLossFn
    q, p, n = triplet
    return TripletLossFn.apply(q, p, n, float(margin))


class ModelFactory:
    """`trainer.py:7,307-310` imports this from `model`; the reference never defined it (SURVEY
    quirk #8).  Only the triplet loss exists in the reference, so that is what it hands out."""

    @staticmethod
    def get_loss_function(loss_type: str = "triplet", margin: float = 1.0):
        if loss_type != "triplet":
            raise ValueError(f"unknown loss_type {loss_type!r}; the reference only implements 'triplet'")
        return lambda triplet: triplet_loss_cosine(triplet, margin=margin)

    @staticmethod
    def create_model(config: Dict, pretrained_embeddings: Optional[np.ndarray] = None) -> TwoTowerModel:
        return TwoTowerModel(config, pretrained_embeddings)
