"""Host-side orchestration of the tower kernels (forward + autograd).

Pipeline of one encode call — the body of `RNNEncoder.forward` (reference
`backend/model.py:48-75`) as stream-ordered launches with no host synchronisation unless
`strict_lengths` asks for the reference's zero-length error:

  ttr_seq_plan        lengths / length-sorted packing plan        (model.py:52-57)
  ttr_embed_gather    packed token matrix X [tokens, E]           (model.py:49)
  per layer:
    ttr_gemm_tf32_bias        gi = X W_ih^T + b_ih, both directions (tcgen05)
    ttr_gru_recurrence_fwd_ws h_t recurrence, per-step outputs, final states
  ttr_proj_l2norm_fwd  cat -> Linear -> F.normalize                (model.py:65-74)
"""
from __future__ import annotations

import os
from typing import List, Optional

import torch

from . import _lib
from .model import _ZERO_LEN_MSG


def _fp16_pipeline_allowed() -> bool:
    """Inference runs gather -> projection -> recurrence with fp16 STORAGE of X, gi and the inter-layer y
    (fp32 accumulation, bias, state) unless TTR_FP32_PIPELINE=1, the fp32 debug GEMM, or a debug flag that
    forces one of the fp32-only recurrence kernels asks otherwise."""
    import ctypes
    if os.environ.get("TTR_FP32_PIPELINE", "0") == "1" or _use_debug_gemm():
        return False
    flags = ctypes.c_int(0)
    _lib.call_nostream("ttr_debug_get_flags", ctypes.byref(flags))
    return (flags.value & (1 | 2 | 1024)) == 0


def _use_debug_gemm() -> bool:
    # test-only switch: route the input projection through the fp32 CUDA-core reference kernel
    return os.environ.get("TTR_DEBUG_FP32_GEMM", "0") == "1"


def input_projection(A: torch.Tensor, W: torch.Tensor, bias: torch.Tensor, out: torch.Tensor, m_bound: int,
                     m_valid: Optional[torch.Tensor]):
    N, K = W.shape
    name = "ttr_debug_gemm_fp32_bias" if _use_debug_gemm() else "ttr_gemm_tf32_bias"
    _lib.call(name, A, W, bias, out, m_bound, m_valid, N, K)


class SeqPlan:
    """Device-resident packing plan of one padded id batch."""

    def __init__(self, ids: torch.Tensor, token_bound: Optional[int] = None):
        B, T = ids.shape
        dev = ids.device
        self.B, self.T = B, T
        self.ids = ids
        self.lengths = torch.empty(B, dtype=torch.int32, device=dev)
        self.order = torch.empty(B, dtype=torch.int32, device=dev)
        self.offsets = torch.empty(B + 1, dtype=torch.int32, device=dev)
        self.status = torch.empty(4, dtype=torch.int32, device=dev)
        _lib.call("ttr_seq_plan", ids, B, T, self.lengths, self.order, self.offsets, self.status)
        # rows allocated for the packed per-token matrices: B*T unless the caller knows the token count
        # (bulk encode: the host tokenised the rows) — rounded up to a 128-row GEMM tile
        self.m_bound = B * T if token_bound is None else min(B * T, (int(token_bound) + 127) // 128 * 128)
        self.total = self.offsets[B:]          # device scalar view: number of packed tokens

    def check_lengths(self):
        st = self.status.cpu()
        if int(st[0]) > 0:
            raise RuntimeError(_ZERO_LEN_MSG)


def _w16_cached(enc, layer: int, W_ih: torch.Tensor) -> torch.Tensor:
    """fp16 copy of a layer's W_ih for the kind::f16 projection, converted once per parameter state: the key is
    the weights' address, torch's version counters of the parameters and of the flat buffer they are views of
    (bumped by in-place torch updates, optimiser steps, load_state_dict) and the owner's `_param_epoch`
    (bumped by FusedClipAdam, whose kernel writes through raw pointers)."""
    owner = enc._owner if enc._owner is not None else enc
    pv = tuple(getattr(enc.rnn, f"weight_ih_l{layer}{sfx}")._version
               for sfx in ([""] + (["_reverse"] if enc.bidirectional else [])))
    key = (W_ih.data_ptr(), W_ih._version, pv, getattr(owner, "_param_epoch", 0))
    cache = enc.__dict__.setdefault("_w16_cache", {})
    hit = cache.get(layer)
    if hit is not None and hit[0] == key:
        return hit[1]
    W16 = torch.empty(W_ih.shape, dtype=torch.float16, device=W_ih.device)
    _lib.call("ttr_f32_to_f16", W_ih.detach().contiguous(), W16, W_ih.numel())
    cache[layer] = (key, W16)
    return W16


def _forward_layers_f16(enc, ids, plan, B, T, H, E, dirs, table, dev) -> torch.Tensor:
    """The inference pipeline with fp16 storage between the kernels (no autograd, no dropout):
    gather -> fp16 X; per layer kind::f16 projection -> fp16 gi; tcgen05 recurrence -> fp16 y / fp32 h_last."""
    Mb = plan.m_bound
    X = torch.empty(Mb, E, dtype=torch.float16, device=dev)
    _lib.call("ttr_embed_gather", ids, B, T, table.detach(), table.shape[0], E, plan.order, plan.offsets, X, 2)
    ws_bytes = int(_lib.load().ttr_gru_fwd_workspace_bytes(B, H, dirs))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    layer_in, h_last = X, None
    for layer in range(enc.num_layers):
        W_ih, b_ih, W_hh, b_hh = enc.layer_weights(layer)
        W16 = _w16_cached(enc, layer, W_ih)
        gi = torch.empty(Mb, dirs * 3 * H, dtype=torch.float16, device=dev)
        _lib.call("ttr_gemm_f16_bias", layer_in, W16, b_ih.detach(), gi, Mb, plan.total, dirs * 3 * H, layer_in.shape[1])
        last = layer == enc.num_layers - 1
        y = None if last else torch.empty(Mb, dirs * H, dtype=torch.float16, device=dev)
        h_last = torch.empty(B, dirs * H, dtype=torch.float32, device=dev)
        _lib.call("ttr_gru_recurrence_fwd_f16", gi, W_hh.detach(), b_hh.detach(), plan.order, plan.offsets, B, H, dirs,
                  y, h_last, ws, ws_bytes)
        del gi
        layer_in = y
    return h_last


def _forward_impl(enc, ids: torch.Tensor, need_grad: bool, training: bool):
    """Runs the kernels; returns (out, ctx dict for backward or None)."""
    if ids.dim() != 2:
        raise ValueError(f"expected int64 ids of shape [B, T], got {tuple(ids.shape)}")
    _lib.require_cuda(ids, "RNNEncoder.forward(x)")
    ids = ids.contiguous()
    if ids.dtype != torch.int64:
        ids = ids.long()
    enc._ensure_flat()
    B, T = ids.shape
    if B == 0 or T == 0:
        raise RuntimeError(_ZERO_LEN_MSG)
    dev = ids.device
    H = enc.hidden_dim
    dirs = 2 if enc.bidirectional else 1
    E = enc.embedding.embedding_dim
    table = enc.embedding.weight
    if table.device != dev:
        raise _lib.TTRError("embedding table and ids are on different devices")

    plan = SeqPlan(ids, getattr(enc, "_token_bound", None) if not need_grad else None)
    if enc.strict_lengths:
        plan.check_lengths()
    Mb = plan.m_bound
    tf32 = not _use_debug_gemm()
    no_dropout = not training or enc.rnn.dropout == 0.0 or enc.num_layers == 1
    if not need_grad and no_dropout and H == 256 and E % 8 == 0 and _fp16_pipeline_allowed():
        h_last = _forward_layers_f16(enc, ids, plan, B, T, H, E, dirs, table, dev)
        layer_ins = ys = saveds = masks = None
    else:
        X = torch.empty(Mb, E, dtype=torch.float32, device=dev)
        _lib.call("ttr_embed_gather", ids, B, T, table.detach(), table.shape[0], E, plan.order, plan.offsets, X,
                  1 if tf32 else 0)

        p_drop = enc.rnn.dropout if training else 0.0
        layer_in = X
        layer_ins: List[torch.Tensor] = []      # GEMM input of each layer (X, then dropped outputs)
        ys: List[Optional[torch.Tensor]] = []   # un-dropped per-step outputs of each layer
        saveds: List[Optional[torch.Tensor]] = []
        masks: List[Optional[torch.Tensor]] = []
        h_last = None
        for layer in range(enc.num_layers):
            W_ih, b_ih, W_hh, b_hh = enc.layer_weights(layer)
            layer_ins.append(layer_in)
            gi = torch.empty(Mb, dirs * 3 * H, dtype=torch.float32, device=dev)
            input_projection(layer_in, W_ih, b_ih, gi, Mb, plan.total)
            last = layer == enc.num_layers - 1
            y = torch.empty(Mb, dirs * H, dtype=torch.float32, device=dev) if (not last or need_grad) else None
            saved = torch.empty(Mb, dirs, 4, H, dtype=torch.float32, device=dev) if need_grad else None
            h_last = torch.empty(B, dirs * H, dtype=torch.float32, device=dev)
            ws_bytes = int(_lib.load().ttr_gru_fwd_workspace_bytes(B, H, dirs))
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev) if ws_bytes else None
            _lib.call("ttr_gru_recurrence_fwd_ws", gi, W_hh, b_hh, plan.order, plan.offsets, B, H, dirs, y, h_last,
                      saved, ws, ws_bytes)
            del gi
            ys.append(y)
            saveds.append(saved)
            mask = None
            if not last:
                layer_in = y
                if p_drop > 0.0:
                    # inter-layer dropout of nn.GRU (model.py:35): scaled keep mask on layer outputs
                    # (seed drawn from torch's CPU generator so torch.manual_seed controls it)
                    seed = int(torch.randint(0, 2 ** 62, (1,)).item())
                    mask = torch.empty_like(y)
                    layer_in = torch.empty_like(y)
                    _lib.call("ttr_dropout", y, y.numel(), float(p_drop), seed, layer_in, mask)
            masks.append(mask)
    out = torch.empty(B, H, dtype=torch.float32, device=dev)
    raw = torch.empty(B, H, dtype=torch.float32, device=dev) if need_grad else None
    if enc.projection is not None:
        Wp = enc._w("projection.weight", (H, dirs * H))
        bp = enc._w("projection.bias", (H,))
    else:
        Wp = bp = None
    _lib.call("ttr_proj_l2norm_fwd", h_last, Wp, bp, B, dirs * H, H, 1 if enc.normalize_output else 0, out, raw)
    enc.last_dropout_masks = masks if training else None
    enc.last_plan = plan
    ctx = None
    if need_grad:
        ctx = dict(plan=plan, layer_ins=layer_ins, ys=ys, saveds=saveds, masks=masks, h_last=h_last, raw=raw)
    return out, ctx


class EncoderFn(torch.autograd.Function):
    """Autograd node of one tower forward.  Parameters are passed as inputs so standard
    `loss.backward()` / optimisers work exactly like with the reference module
    (backend/main.py:254-259)."""

    @staticmethod
    def forward(ctx, enc, ids, *params):
        out, saved = _forward_impl(enc, ids, need_grad=True, training=enc.training)
        ctx.enc = enc
        ctx.saved = saved
        ctx.n_params = len(params)
        return out

    @staticmethod
    def backward(ctx, d_out):
        from .towers_bwd import encoder_backward
        grads = encoder_backward(ctx.enc, ctx.saved, d_out.contiguous())
        ctx.saved = None
        return (None, None) + tuple(grads)


def encoder_params(enc) -> List[torch.nn.Parameter]:
    """Parameters handed to autograd, in `encoder_backward`'s gradient order."""
    from .model import _flat_order
    ps = [p for _, p in _flat_order(enc)]
    if enc.embedding.weight.requires_grad:
        ps.append(enc.embedding.weight)
    return ps


def encoder_forward(enc, ids: torch.Tensor) -> torch.Tensor:
    params = encoder_params(enc)
    need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in params)
    if not need_grad:
        out, _ = _forward_impl(enc, ids, need_grad=False, training=enc.training)
        return out
    return EncoderFn.apply(enc, ids, *params)


class TripletLossFn(torch.autograd.Function):
    """`triplet_loss_cosine` (backend/model.py:109-114) forward/backward kernels."""

    @staticmethod
    def forward(ctx, q, p, n, margin):
        for t in (q, p, n):
            _lib.require_cuda(t, "triplet_loss_cosine")
        q, p, n = q.contiguous().float(), p.contiguous().float(), n.contiguous().float()
        B, H = q.shape
        loss = torch.empty(1, dtype=torch.float32, device=q.device)
        _lib.call("ttr_triplet_fwd", q, p, n, B, H, float(margin), loss)
        ctx.save_for_backward(q, p, n)
        ctx.margin = float(margin)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, dloss):
        q, p, n = ctx.saved_tensors
        B, H = q.shape
        dq, dp, dn = torch.empty_like(q), torch.empty_like(p), torch.empty_like(n)
        dl = dloss.reshape(1).contiguous().float()
        _lib.call("ttr_triplet_bwd", q, p, n, B, H, ctx.margin, dl, dq, dp, dn)
        return dq, dp, dn, None
