"""Drop-in `TwoTowerTrainer` / `TrainerFactory` — reference `backend/trainer.py:12-317`.

The reference file cannot be imported (it needs `ModelFactory` and `utils.clean_memory`,
neither of which exists — SURVEY quirk #8); the numerical contract of a step is the live loop
in `backend/main.py:244-259`.  This module keeps the reference's class, method and metric
names, and runs every step on the sm_100a kernels:

  3 encodes -> cosine triplet loss -> backward -> [all-reduce] -> fused clip + Adam

`clip_max_norm` selects between the two reference variants: 1.0 reproduces `main.py:257`
(the live loop), None the un-clipped `trainer.py:101-110`.  W&B logging is reduced to an
optional callback — there is no network here and logging is not on the hot path.
"""
from __future__ import annotations

import gc
import time
from typing import Callable, Dict, List, Optional

import numpy as np
import torch

from . import _lib
from .index import search_topk
from .model import ModelFactory, TwoTowerModel
from .optim import FusedClipAdam


def clean_memory():
    """`backend/main.py:68-74`."""
    gc.collect()
    if torch.cuda.is_available():
        torch.cuda.empty_cache()


class TwoTowerTrainer:
    def __init__(self, model: TwoTowerModel, optimizer, loss_function, device: torch.device, config: Dict,
                 log_fn: Optional[Callable[[dict], None]] = None):
        self.model = model
        self.optimizer = optimizer
        self.loss_function = loss_function
        self.device = device
        self.config = config
        self.train_losses: List[float] = []
        self.val_losses: List[float] = []
        self.best_val_loss = float("inf")
        self.log_fn = log_fn or (lambda d: None)
        # The three tower passes of a step (query, positive, negative) are independent until the loss.  Each is a chain
        # of latency-bound recurrence kernels that leave SMs, issue slots and HBM idle, so they run on three streams:
        # one tower's projection / weight-gradient GEMMs fill the gaps of another tower's recurrence.  autograd runs
        # every backward node on the stream of its forward, so the backward passes overlap the same way.
        # 1 = one stream (everything in program order).
        self.tower_streams = int(config.get("TOWER_STREAMS", 3))
        self._lanes: List[torch.cuda.Stream] = []

    # ------------------------------------------------------------------ metrics
    def compute_batch_metrics(self, q_vec, pos_vec, neg_vec) -> Dict[str, float]:
        """`trainer.py:38-55` in one kernel and one device->host read instead of five."""
        with torch.no_grad():
            B, H = q_vec.shape
            out = torch.empty(5, dtype=torch.float32, device=q_vec.device)
            _lib.call("ttr_batch_metrics", q_vec.detach().contiguous(), pos_vec.detach().contiguous(),
                      neg_vec.detach().contiguous(), B, H, out)
            acc, gap, mag, pos, neg = out.cpu().tolist()
        return {"accuracy": acc, "similarity_gap": gap, "magnitude": mag,
                "pos_similarity": pos, "neg_similarity": neg}

    def compute_recall_metrics(self, query_embeddings, doc_embeddings, k_values=[5, 10]) -> Dict[str, float]:
        """`trainer.py:57-78`: is document i among query i's top-k?  Fused score+top-k kernel when the
        embeddings are 256-wide, otherwise a plain matmul (metric bookkeeping, not the hot path)."""
        with torch.no_grad():
            B = query_embeddings.size(0)
            kmax = min(max(k_values), doc_embeddings.size(0))
            if query_embeddings.shape[1] == 256 and kmax <= 64:
                _, top = search_topk(query_embeddings.detach(), doc_embeddings.detach().contiguous(), kmax)
            else:
                _, top = torch.topk(query_embeddings @ doc_embeddings.t(), k=kmax, dim=1)
            target = torch.arange(B, device=top.device).unsqueeze(1)
            return {f"recall_at_{k}": float((top[:, :k] == target).any(dim=1).float().mean()) for k in k_values}

    # ------------------------------------------------------------------ one step
    def train_step(self, query_batch, pos_batch, neg_batch):
        """One optimisation step — `backend/main.py:244-259`.  Returns (loss tensor, q, p, n)."""
        query_batch = query_batch.to(self.device, non_blocking=True)
        pos_batch = pos_batch.to(self.device, non_blocking=True)
        neg_batch = neg_batch.to(self.device, non_blocking=True)
        self.optimizer.zero_grad()
        if self.tower_streams > 1 and query_batch.is_cuda:
            q_vec, pos_vec, neg_vec = self._encode_on_lanes(query_batch, pos_batch, neg_batch)
        else:
            q_vec = self.model.encode_query(query_batch)
            pos_vec = self.model.encode_document(pos_batch)
            neg_vec = self.model.encode_document(neg_batch)
        loss = self.loss_function((q_vec, pos_vec, neg_vec))
        loss.backward()
        self.optimizer.step()
        return loss.detach(), q_vec.detach(), pos_vec.detach(), neg_vec.detach()

    def _encode_on_lanes(self, query_batch, pos_batch, neg_batch):
        dev = query_batch.device
        main = torch.cuda.current_stream(dev)
        n = min(self.tower_streams, 3)
        if len(self._lanes) != n or self._lanes[0].device != dev:
            self._lanes = [torch.cuda.Stream(device=dev) for _ in range(n)]
            # the parameters' AccumulateGrad nodes live on the stream the model was built on; gradients now arrive from
            # the lanes, which autograd synchronises correctly but reports on every backward
            quiet = getattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch", None)
            if quiet is not None:
                quiet(False)
        self.model._ensure_flat()                 # lazily rebuilt buffers must exist before the lanes fork
        for s in self._lanes:
            s.wait_stream(main)
        outs = []
        jobs = ((self.model.encode_query, query_batch), (self.model.encode_document, pos_batch),
                (self.model.encode_document, neg_batch))
        for i, (fn, ids) in enumerate(jobs):
            s = self._lanes[i % n]
            with torch.cuda.stream(s):
                e = fn(ids)
            ids.record_stream(s)
            e.record_stream(main)
            outs.append(e)
        for s in self._lanes:
            main.wait_stream(s)
        return outs

    def train_epoch(self, train_loader, val_loader, epoch):
        self.model.train()
        total_loss = torch.zeros((), device=self.device)
        totals = {"accuracy": 0.0, "similarity_gap": 0.0, "magnitude": 0.0}
        num_batches = 0
        for query_batch, pos_batch, neg_batch in train_loader:
            loss, q, p, n = self.train_step(query_batch, pos_batch, neg_batch)
            bm = self.compute_batch_metrics(q, p, n)
            total_loss += loss                       # stays on the device: no per-step .item() sync
            for k in totals:
                totals[k] += bm[k]
            num_batches += 1
            if num_batches % 50 == 0:
                self.log_fn({"batch_loss": float(loss), "batch_accuracy": bm["accuracy"],
                             "batch_similarity_gap": bm["similarity_gap"], "batch_magnitude": bm["magnitude"],
                             "batch": num_batches, "epoch": epoch + 1})
            if num_batches % 200 == 0 and val_loader is not None:
                self.model.eval()
                rm = self.quick_recall_check(val_loader)
                self.log_fn({**{f"batch_{k}": v for k, v in rm.items()}, "batch": num_batches, "epoch": epoch + 1})
                self.model.train()
        n = max(num_batches, 1)
        return float(total_loss) / n, {k: v / n for k, v in totals.items()}

    def validate_epoch(self, val_loader, epoch):
        self.model.eval()
        total_loss, totals, num_batches = 0.0, {"accuracy": 0.0, "similarity_gap": 0.0, "magnitude": 0.0}, 0
        with torch.no_grad():
            for query_batch, pos_batch, neg_batch in val_loader:
                q = self.model.encode_query(query_batch.to(self.device, non_blocking=True))
                p = self.model.encode_document(pos_batch.to(self.device, non_blocking=True))
                n = self.model.encode_document(neg_batch.to(self.device, non_blocking=True))
                total_loss += float(self.loss_function((q, p, n)))
                bm = self.compute_batch_metrics(q, p, n)
                for k in totals:
                    totals[k] += bm[k]
                num_batches += 1
        n = max(num_batches, 1)
        return total_loss / n, {k: v / n for k, v in totals.items()}

    def quick_recall_check(self, val_loader, max_batches=3):
        all_metrics = []
        with torch.no_grad():
            for batch_idx, (query_batch, pos_batch, neg_batch) in enumerate(val_loader):
                if batch_idx >= max_batches:
                    break
                q = self.model.encode_query(query_batch.to(self.device, non_blocking=True))
                p = self.model.encode_document(pos_batch.to(self.device, non_blocking=True))
                n = self.model.encode_document(neg_batch.to(self.device, non_blocking=True))
                all_metrics.append(self.compute_recall_metrics(q, torch.cat([p, n], dim=0)))
        if not all_metrics:
            return {}
        return {k: float(np.mean([m[k] for m in all_metrics])) for k in all_metrics[0]}

    def train(self, train_loader, val_loader=None, epochs=None):
        if epochs is None:
            epochs = self.config.get("EPOCHS", 10)
        start = time.time()
        for epoch in range(epochs):
            train_loss, train_metrics = self.train_epoch(train_loader, val_loader, epoch)
            self.train_losses.append(train_loss)
            log = {"epoch": epoch + 1, "train_loss": train_loss,
                   **{f"train_{k}": v for k, v in train_metrics.items()}}
            if val_loader is not None:
                val_loss, val_metrics = self.validate_epoch(val_loader, epoch)
                self.val_losses.append(val_loss)
                self.best_val_loss = min(self.best_val_loss, val_loss)
                rm = self.quick_recall_check(val_loader, max_batches=5)
                log.update({"val_loss": val_loss, **{f"val_{k}": v for k, v in val_metrics.items()},
                            **{f"epoch_{k}": v for k, v in rm.items()}})
            self.log_fn(log)
            clean_memory()
        self.total_time = time.time() - start
        return {"train_losses": self.train_losses, "val_losses": self.val_losses,
                "best_val_loss": self.best_val_loss}


class TrainerFactory:
    """`trainer.py:298-317`: Adam(lr=config['LR'] or 1e-3) + triplet loss(margin=config['MARGIN'] or 1.0)."""

    @staticmethod
    def create_trainer(config: Dict, model: TwoTowerModel, device: torch.device, fused: bool = True,
                       clip_max_norm: Optional[float] = None, process_group=None) -> TwoTowerTrainer:
        lr = config.get("LR", 0.001)
        if fused:
            optimizer = FusedClipAdam(model, lr=lr, max_norm=clip_max_norm, process_group=process_group)
        else:
            optimizer = torch.optim.Adam(model.parameters(), lr=lr)
        loss_function = ModelFactory.get_loss_function(loss_type=config.get("LOSS_TYPE", "triplet"),
                                                       margin=config.get("MARGIN", 1.0))
        return TwoTowerTrainer(model=model, optimizer=optimizer, loss_function=loss_function, device=device,
                               config=config)
