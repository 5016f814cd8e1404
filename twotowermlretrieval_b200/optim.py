"""Fused clip_grad_norm_ + Adam over the model's flat parameter bucket, with the
data-parallel gradient all-reduce in front of it.

Replaces `torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=1.0)` +
`torch.optim.Adam.step()` of the live training loop (reference `backend/main.py:222,257-259`).
The reference is single-process; the all-reduce (one NCCL call over the 16 MB bucket, SUM then
1/world inside the kernel) is the data-parallel extension described in SURVEY.md §8(e).
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib


class FusedClipAdam:
    """Duck-types the parts of `torch.optim.Optimizer` the reference loop touches
    (`zero_grad`, `step`, `param_groups`, `state_dict`).  Only the GRU/projection parameters
    that live in the flat bucket are updated; a trainable embedding table (no pretrained
    matrix) is delegated to a regular torch Adam so semantics stay those of the reference."""

    def __init__(self, model, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 max_norm: Optional[float] = 1.0, process_group=None):
        self.model = model
        self.param_groups = [dict(lr=lr, betas=betas, eps=eps, max_norm=max_norm)]
        self.step_count = 0
        self.group = process_group
        self._m = self._v = self._ws = None
        self.last_grad_norm: Optional[torch.Tensor] = None
        extra = [p for n, p in model.named_parameters() if n.endswith("embedding.weight") and p.requires_grad]
        self._extra_params = extra
        self._extra = torch.optim.Adam(extra, lr=lr, betas=betas, eps=eps) if extra else None

    def _state(self):
        flat = self.model.flat_params()
        if self._m is None or self._m.numel() != flat.numel() or self._m.device != flat.device:
            self._m = torch.zeros_like(flat)
            self._v = torch.zeros_like(flat)
            self._ws = torch.empty(1024, dtype=torch.float32, device=flat.device)
            self.last_grad_norm = torch.zeros(1, dtype=torch.float32, device=flat.device)
        return flat

    def zero_grad(self, set_to_none: bool = False):
        self.model.flat_grads().zero_()
        if self._extra is not None:
            self._extra.zero_grad(set_to_none=True)

    def step(self):
        import torch.distributed as dist
        flat = self._state()
        grads = self.model.flat_grads()
        g = self.param_groups[0]
        world = 1
        if dist.is_available() and dist.is_initialized():
            world = dist.get_world_size(self.group)
            if world > 1:
                dist.all_reduce(grads, op=dist.ReduceOp.SUM, group=self.group)
                for p in self._extra_params:
                    if p.grad is not None:
                        dist.all_reduce(p.grad, op=dist.ReduceOp.SUM, group=self.group)
                        p.grad.div_(world)
        self.step_count += 1
        max_norm = g["max_norm"]
        if self._extra_params and max_norm is not None and max_norm > 0:
            # the reference clips over ALL parameters jointly; with a trainable table fall back to
            # torch's clip so the joint norm is exact, then run Adam without clipping
            if world > 1:
                grads.div_(world)
            torch.nn.utils.clip_grad_norm_(list(self.model.parameters()), max_norm)
            scale, max_norm = 1.0, -1.0
        else:
            scale = 1.0 / world
        _lib.call("ttr_clip_adam", flat, grads, self._m, self._v, flat.numel(), float(scale),
                  float(max_norm if max_norm is not None else -1.0), float(g["lr"]), float(g["betas"][0]),
                  float(g["betas"][1]), float(g["eps"]), int(self.step_count), self.last_grad_norm, self._ws)
        # the kernel updated the parameters through raw pointers: invalidate caches keyed on the parameter state
        self.model._param_epoch = getattr(self.model, "_param_epoch", 0) + 1
        if self._extra is not None:
            self._extra.step()

    def state_dict(self):
        return {"step": self.step_count, "exp_avg": self._m, "exp_avg_sq": self._v,
                "param_groups": self.param_groups,
                # moments of a trainable embedding table (delegated to torch.optim.Adam)
                "embedding_adam": self._extra.state_dict() if self._extra is not None else None}

    def load_state_dict(self, sd):
        self.step_count = int(sd["step"])
        self._state()
        if sd.get("exp_avg") is not None:
            self._m.copy_(sd["exp_avg"])
            self._v.copy_(sd["exp_avg_sq"])
        if self._extra is not None and sd.get("embedding_adam") is not None:
            self._extra.load_state_dict(sd["embedding_adam"])
