"""Backward of one tower forward: the autograd of `RNNEncoder.forward`
(reference `backend/model.py:48-75`) reached from `loss.backward()` (`backend/main.py:254`).

  ttr_l2norm_bwd            through F.normalize
  gemm_nn / gemm_tn / colsum through the projection Linear
  per layer, top down:
    ttr_gru_recurrence_bwd_ws  BPTT -> d(gi), d(gh) per token (tcgen05 split-K for H = 256)
    ttr_gru_whh_grad        dW_hh = d(gh)^T h_prev
    ttr_gemm_tn_fp32        dW_ih = d(gi)^T layer_input
    ttr_colsum              db_ih, db_hh
    ttr_gemm_nn_fp32        d(layer_input) = d(gi) W_ih   (times the dropout mask)
  ttr_embed_scatter_grad    only when the table is trainable
"""
from __future__ import annotations

from typing import List, Optional

import torch

from . import _lib


def encoder_backward(enc, ctx: dict, d_out: torch.Tensor) -> List[Optional[torch.Tensor]]:
    """Returns gradients in `towers.encoder_params(enc)` order."""
    from .model import _flat_order
    plan = ctx["plan"]
    B, Mb = plan.B, plan.m_bound
    dev = d_out.device
    H = enc.hidden_dim
    dirs = 2 if enc.bidirectional else 1
    G = dirs * 3 * H
    E = enc.embedding.embedding_dim
    named = _flat_order(enc)
    # one flat gradient buffer with the same relative layout as the parameters
    sizes = [p.numel() for _, p in named]
    flat = torch.zeros(sum(sizes), dtype=torch.float32, device=dev)
    views, o = {}, 0
    for (name, p), n in zip(named, sizes):
        views[name] = flat[o:o + n]
        o += n

    def gview(kind: str, layer: int):
        """Contiguous gradient region of both directions of `kind` at `layer`."""
        first = f"rnn.{kind}_l{layer}"
        start = views[first].data_ptr() - flat.data_ptr()
        n = views[first].numel() * dirs
        s = start // 4
        return flat[s:s + n]

    d_out = d_out.float().contiguous()
    d_raw = torch.empty(B, H, dtype=torch.float32, device=dev)
    _lib.call("ttr_l2norm_bwd", d_out, ctx["raw"], B, H, 1 if enc.normalize_output else 0, d_raw)
    h_last = ctx["h_last"]
    if enc.projection is not None:
        Wp = enc._w("projection.weight", (H, dirs * H))
        d_hcat = torch.empty(B, dirs * H, dtype=torch.float32, device=dev)
        _lib.call("ttr_gemm_nn_fp32", d_raw, Wp, d_hcat, B, None, H, dirs * H, 0)
        _lib.call("ttr_gemm_tn_fp32", d_raw, h_last, views["projection.weight"], B, None, H, dirs * H, 0)
        _lib.call("ttr_colsum", d_raw, B, None, H, views["projection.bias"], 0)
    else:
        d_hcat = d_raw

    dy = None
    dh_last = d_hcat
    d_table = None
    for layer in range(enc.num_layers - 1, -1, -1):
        W_ih, _, W_hh, _ = enc.layer_weights(layer)
        in_dim = W_ih.shape[1]
        dgi = torch.empty(Mb, G, dtype=torch.float32, device=dev)
        dgh = torch.empty(Mb, G, dtype=torch.float32, device=dev)
        ws_bytes = int(_lib.load().ttr_gru_bwd_workspace_bytes(B, H, dirs))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev) if ws_bytes else None
        _lib.call("ttr_gru_recurrence_bwd_ws", dy, dh_last, ctx["ys"][layer], ctx["saveds"][layer], W_hh,
                  plan.order, plan.offsets, B, H, dirs, dgi, dgh, ws, ws_bytes, Mb, plan.total,
                  gview("bias_ih", layer), gview("bias_hh", layer))       # bias gradients accumulate into the zeroed views
        # weight gradients on the tensor cores (tf32, token dimension reduced with split-K); the
        # operands' rows between the valid token count and the next multiple of 32 must be zero
        layer_in = ctx["layer_ins"][layer]
        for t, ld in ((dgi, G), (dgh, G), (layer_in, in_dim)):
            _lib.call("ttr_zero_tail_rows", t, Mb, plan.total, ld)
        hprev = torch.empty(Mb, dirs * H, dtype=torch.float32, device=dev)
        _lib.call("ttr_gru_whh_grad", dgh, ctx["ys"][layer], plan.offsets, B, H, dirs, Mb, hprev,
                  gview("weight_hh", layer), 0)
        del hprev
        _lib.call("ttr_gemm_tn_tf32", dgi, G, layer_in, in_dim, gview("weight_ih", layer), in_dim, Mb, plan.total,
                  G, in_dim, 0)
        del dgh
        if layer > 0:
            # dX = dG W_ih through the tcgen05 NT kernel: B operand = W_ih^T [in, G] (3 MB transpose)
            dx = torch.empty(Mb, in_dim, dtype=torch.float32, device=dev)
            _lib.call("ttr_gemm_tf32_bias", dgi, W_ih.t().contiguous(), None, dx, Mb, plan.total, in_dim, G)
            mask = ctx["masks"][layer - 1]
            if mask is not None:
                dx.mul_(mask)
            dy, dh_last = dx, None
        elif enc.embedding.weight.requires_grad:
            dx = torch.empty(Mb, E, dtype=torch.float32, device=dev)
            _lib.call("ttr_gemm_nn_fp32", dgi, W_ih, dx, Mb, plan.total, G, E, 0)
            table = enc.embedding.weight
            d_table = torch.zeros_like(table)
            _lib.call("ttr_embed_scatter_grad", plan.ids, B, plan.T, table.shape[0], E, plan.order, plan.offsets,
                      dx, d_table)
        del dgi

    grads: List[Optional[torch.Tensor]] = [views[name].view(p.shape) for name, p in named]
    if enc.embedding.weight.requires_grad:
        grads.append(d_table)
    return grads
