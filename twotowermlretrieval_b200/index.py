"""Exact dense retrieval + hybrid rerank on the GPU (additive API).

Replaces the reference's `torch.matmul(q, D.t())` + `torch.topk` call sites
(`backend/evaluators.py:185-186`, `backend/trainer.py:62-65`), the ChromaDB ANN query of
the frontend (`frontend/main.py:153-156`) and the Python rerank loop
(`frontend/main.py:158-198`).  Row shards of the document matrix live on different GPUs;
the only exchange is an all-gather of the per-rank top-k candidate lists.
"""
from __future__ import annotations

import threading
from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib

KMAX = 64          # TOPK_KMAX of the kernels (csrc/topk_common.cuh)
SPACE_L2 = 0       # Chroma default: semantic = 1 - ||q-d||^2 = 2cos - 1 (frontend/main.py:162)
SPACE_COSINE = 1   # semantic = cos


def _space_code(space) -> int:
    if space in (SPACE_L2, "l2"):
        return SPACE_L2
    if space in (SPACE_COSINE, "cosine"):
        return SPACE_COSINE
    raise ValueError(f"unknown space {space!r} (use 'l2' or 'cosine')")


class _Workspace:
    """Scratch of the scorer (bounds, candidate lists, grid-barrier counters) for ONE (device, stream): launches on
    different streams never share scratch, so concurrent searches are safe (re-entrant per CUDA stream).
    `lock` only keeps two host threads that enqueue on the SAME stream (the reference serves `/search` from a thread
    pool, frontend/main.py:102-103) from interleaving the launches of their calls; nothing is synchronised."""
    __slots__ = ("buf", "lock")

    def __init__(self):
        self.buf = None
        self.lock = threading.Lock()


_WORKSPACES: dict = {}
_WS_GUARD = threading.Lock()


def _workspace(device, kind: str = "score") -> _Workspace:
    dev = device.index if device.index is not None else torch.cuda.current_device()
    key = (kind, dev, torch.cuda.current_stream(dev).cuda_stream)
    ws = _WORKSPACES.get(key)
    if ws is None:
        with _WS_GUARD:
            ws = _WORKSPACES.setdefault(key, _Workspace())
    return ws


def search_topk(Q: torch.Tensor, docs: torch.Tensor, k: int = 50, row_offset: int = 0,
                out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Exact top-k of Q @ docs.T per query, without materialising the score matrix.

    Q fp32 [B, 256], docs fp32 [N, 256] (both CUDA, contiguous).  Returns (scores fp32
    [B, k] descending, idx int64 [B, k] + row_offset); ties: lower index first.  k is
    clamped to N like the reference's evaluators do implicitly (`k=max(top_k)` on small
    candidate sets would raise in torch; here missing slots carry -inf / -1)."""
    _lib.require_cuda(Q, "search_topk(Q)")
    _lib.require_cuda(docs, "search_topk(docs)")
    if not 1 <= k <= KMAX:
        raise ValueError(f"search_topk: k={k} outside [1, {KMAX}] (the fused kernels keep at most {KMAX} "
                         "candidates per query; the reference's own call sites use k <= 50)")
    if Q.dim() == 1:
        Q = Q.unsqueeze(0)
    Q = Q.contiguous().float()
    if not docs.is_contiguous() or docs.dtype != torch.float32:
        raise _lib.TTRError("search_topk: docs must be a contiguous float32 [N, D] matrix")
    B, D = Q.shape
    N = docs.shape[0]
    lib = _lib.load()
    nbytes = lib.ttr_score_topk_workspace_bytes(B, N, k)
    if out is None:
        scores = torch.empty(B, k, dtype=torch.float32, device=Q.device)
        idx = torch.empty(B, k, dtype=torch.int64, device=Q.device)
    else:
        scores, idx = out
    ws = _workspace(Q.device)
    with ws.lock:
        if ws.buf is None or ws.buf.numel() < nbytes:
            ws.buf = torch.empty(int(nbytes), dtype=torch.uint8, device=Q.device)
        _lib.call("ttr_score_topk", Q, B, docs, N, D, k, int(row_offset), scores, idx, ws.buf, ws.buf.numel())
    return scores, idx


def blend_topk(q: torch.Tensor, q_norm: float, docs: torch.Tensor, csr: "CsrF64", q_idx: torch.Tensor,
               q_val: torch.Tensor, alpha: float, k: int):
    """Corpus-wide alpha*dense_cos + (1-alpha)*tfidf_cos and its top-k in one pass (`ttr_blend_topk`):
    `SimpleHybridRetriever.search` (backend/simple_hybrid.py:45-66) and, with alpha = 0, the keyword branch of
    `/search` (frontend/main.py:119-147).  Returns (scores fp64 [k] descending, local row ids int64 [k])."""
    dev = docs.device
    N, D = docs.shape
    nbytes = int(_lib.load().ttr_blend_topk_workspace_bytes(k))
    out_s = torch.empty(k, dtype=torch.float64, device=dev)
    out_i = torch.empty(k, dtype=torch.int64, device=dev)
    ws = _workspace(dev, kind="blend")
    with ws.lock:
        if ws.buf is None or ws.buf.numel() < nbytes:
            ws.buf = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        _lib.call("ttr_blend_topk", q, float(q_norm), docs, N, D, csr.indptr, csr.indices, csr.data, q_idx, q_val,
                  int(q_idx.numel()), float(alpha), k, out_s, out_i, None, ws.buf)
    return out_s, out_i


def topk_merge(cand_scores: torch.Tensor, cand_idx: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Merge P candidate lists per query: cand_* [P, B, kin] -> (scores [B,k], idx [B,k])."""
    P, B, kin = cand_scores.shape
    cand_scores = cand_scores.contiguous()
    cand_idx = cand_idx.contiguous()
    scores = torch.empty(B, k, dtype=torch.float32, device=cand_scores.device)
    idx = torch.empty(B, k, dtype=torch.int64, device=cand_scores.device)
    _lib.call("ttr_topk_merge", cand_scores, cand_idx, P, B, kin, k, scores, idx)
    return scores, idx


@dataclass
class CsrF64:
    """L2-normalised TF-IDF rows on the device (sklearn `TfidfVectorizer` output layout,
    `backend/main.py:142-143`): indptr int64 [rows+1], indices int32 (sorted per row), data fp64."""
    indptr: torch.Tensor
    indices: torch.Tensor
    data: torch.Tensor
    rows: int
    row_offset: int = 0

    @staticmethod
    def from_arrays(indptr, indices, data, device, row_offset: int = 0) -> "CsrF64":
        ip = torch.as_tensor(np.asarray(indptr, dtype=np.int64), device=device)
        ix = torch.as_tensor(np.asarray(indices, dtype=np.int32), device=device)
        dv = torch.as_tensor(np.asarray(data, dtype=np.float64), device=device)
        return CsrF64(ip, ix, dv, int(ip.numel() - 1), int(row_offset))

    @staticmethod
    def from_scipy(mat, device, row_offset: int = 0) -> "CsrF64":
        mat = mat.tocsr()
        mat.sort_indices()
        return CsrF64.from_arrays(mat.indptr, mat.indices, mat.data, device, row_offset)

    def row_slice(self, lo: int, hi: int) -> "CsrF64":
        """Rows [lo, hi) as an independent CSR (host-side shard split)."""
        ip = self.indptr[lo:hi + 1]
        base = int(ip[0])
        end = int(ip[-1])
        return CsrF64((ip - base).contiguous(), self.indices[base:end].contiguous(),
                      self.data[base:end].contiguous(), hi - lo, self.row_offset + lo)


def tfidf_candidates(cand_idx: torch.Tensor, docs_csr: CsrF64, q_csr: CsrF64,
                     out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """TF-IDF cosine of every candidate this shard owns (0 elsewhere): fp64 [B, kc]."""
    B, kc = cand_idx.shape
    if out is None:
        out = torch.empty(B, kc, dtype=torch.float64, device=cand_idx.device)
    _lib.call("ttr_tfidf_candidates", cand_idx.contiguous(), B, kc, docs_csr.row_offset, docs_csr.rows,
              docs_csr.indptr, docs_csr.indices, docs_csr.data, q_csr.indptr, q_csr.indices, q_csr.data, out)
    return out


def hybrid_rerank(cand_idx: torch.Tensor, cand_cos: torch.Tensor, alpha: float,
                  docs_csr: Optional[CsrF64] = None, q_csr: Optional[CsrF64] = None,
                  tfidf: Optional[torch.Tensor] = None, space="l2", top_n: int = 10,
                  q_sqnorm: Optional[torch.Tensor] = None, d_sqnorm: Optional[torch.Tensor] = None):
    """The `/search` rerank (frontend/main.py:158-198) for a batch of queries.

    cand_idx int64 [B, kc] (global document ids, dense-rank order), cand_cos fp32 [B, kc].
    Either (docs_csr, q_csr) or precomputed `tfidf` fp64 [B, kc] must be given.  `q_sqnorm` fp64 [B] /
    `d_sqnorm` fp64 [B, kc]: squared norms for the general squared-L2 semantic of space 'l2' (zero-vector
    query of a token-less string, un-normalised model); omitted = unit vectors (2 cos - 1).
    Returns dict(final, semantic, tfidf fp64 [B, top_n], pos int32 [B, top_n], idx int64 [B, top_n])."""
    B, kc = cand_idx.shape
    dev = cand_idx.device
    cand_idx = cand_idx.contiguous()
    cand_cos = cand_cos.contiguous().float()
    fin = torch.empty(B, top_n, dtype=torch.float64, device=dev)
    sem = torch.empty_like(fin)
    tf = torch.empty_like(fin)
    pos = torch.empty(B, top_n, dtype=torch.int32, device=dev)
    if q_sqnorm is not None:
        q_sqnorm = q_sqnorm.to(device=dev, dtype=torch.float64).contiguous()
    if d_sqnorm is not None:
        d_sqnorm = d_sqnorm.to(device=dev, dtype=torch.float64).contiguous()
    if tfidf is not None:
        _lib.call("ttr_hybrid_rerank", cand_idx, cand_cos, B, kc, 0, None, None, None, None, None, None,
                  tfidf.contiguous(), q_sqnorm, d_sqnorm, float(alpha), _space_code(space), top_n, fin, sem, tf, pos)
    else:
        if docs_csr is None or q_csr is None:
            raise ValueError("hybrid_rerank needs (docs_csr, q_csr) or tfidf")
        _lib.call("ttr_hybrid_rerank", cand_idx, cand_cos, B, kc, docs_csr.row_offset, docs_csr.indptr,
                  docs_csr.indices, docs_csr.data, q_csr.indptr, q_csr.indices, q_csr.data, None, q_sqnorm, d_sqnorm,
                  float(alpha), _space_code(space), top_n, fin, sem, tf, pos)
    idx = torch.gather(cand_idx, 1, pos.clamp_min(0).long())
    return {"final": fin, "semantic": sem, "tfidf": tf, "pos": pos, "idx": idx}


def shard_bounds(n_rows: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous row shard of rank: rows [r*ceil(N/R), min(N, (r+1)*ceil(N/R)))  (SURVEY §8e)."""
    per = -(-n_rows // world)
    lo = min(n_rows, rank * per)
    return lo, min(n_rows, lo + per)


class PendingSearch:
    """Result of a search whose cross-rank merge runs on the exchange's side stream.  `result()` makes the caller's
    current stream wait for it and returns the tensors; until then the caller's stream is free to start the next
    scan (`ShardedIndex.search_deferred`)."""

    def __init__(self, tensors, event=None):
        self._tensors, self._event = tensors, event

    def result(self):
        if self._event is not None:
            cur = torch.cuda.current_stream(self._tensors[0].device)
            cur.wait_event(self._event)
            for t in self._tensors:
                if t is not None:
                    t.record_stream(cur)
            self._event = None
        return self._tensors


class _PeerExchange:
    """Symmetric-memory candidate buffers: every rank writes its local [B, k] top-k (scores, global
    ids, optional TF-IDF payload) into its own buffer; ONE kernel per step then signals the peers,
    waits for their signals (release/acquire flags in the same symmetric allocation) and merges all
    R lists in place over NVLink (ttr_topk_exchange_merge) — no all-gather, no separate barrier launch.

    The merge kernel runs on a SIDE STREAM behind the scan of its step, and THREE list buffers rotate: the scan
    of step s writes buffer s mod 3, last read by the peers' merges of step s - 3.  Those are finished once this
    rank's merge of step s - 2 has completed (it waited for every peer's flag of step s - 2, which a peer raises in
    its merge kernel of that step, i.e. after its merge of step s - 3 on the same stream) — so the scan of step s
    only waits for that event, and a rank whose peers are late runs one scan ahead instead of spinning in the merge
    (`search_deferred`; at 8 GPUs the synchronous step spent 41 us behind a 231 us scan in exchange + rank skew)."""

    FLAG_BYTES = 256
    NBUF = 3

    def __init__(self, group, device, max_b: int, k: int):
        import ctypes
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        self.ctypes = ctypes
        self.k, self.max_b = k, max_b
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        n = max_b * k
        self.off_s, self.off_i, self.off_t = 0, n * 4, n * 12
        self.stride = (n * 20 + 255) // 256 * 256
        self.off_flags = self.NBUF * self.stride
        self.buf = symm.empty(self.NBUF * self.stride + self.FLAG_BYTES, dtype=torch.uint8, device=device)
        grp = group if group is not None else dist.group.WORLD
        self.hdl = symm.rendezvous(self.buf, grp)
        self.ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        self.buf[self.off_flags:].zero_()
        torch.cuda.current_stream(device).synchronize()
        self.hdl.barrier(channel=0)            # every rank's flags are zero before anyone signals step 1
        self.step = 0
        arr = ctypes.c_uint64 * self.world
        self._arr = arr
        self._flags = arr(*[a + self.off_flags for a in self.ptrs])
        self.merge_stream = torch.cuda.Stream(device=device)
        self._merged = {}                      # step -> event recorded behind that step's merge kernel

    def views(self, B: int):
        """Views of the buffer the NEXT step's scan writes; the caller's stream first waits until the peers can no
        longer be reading it (this rank's merge of two steps ago has completed)."""
        s = self.step + 1
        ev = self._merged.get(s - 2)
        if ev is not None:
            torch.cuda.current_stream(self.buf.device).wait_event(ev)
        base = (s % self.NBUF) * self.stride
        n = B * self.k
        s = self.buf[base + self.off_s: base + self.off_s + n * 4].view(torch.float32).view(B, self.k)
        i = self.buf[base + self.off_i: base + self.off_i + n * 8].view(torch.int64).view(B, self.k)
        t = self.buf[base + self.off_t: base + self.off_t + n * 8].view(torch.float64).view(B, self.k)
        return s, i, t

    def merge(self, B: int, with_tfidf: bool, deferred: bool = False):
        """One kernel on the side stream, behind everything the caller's stream has enqueued so far (the scan that
        filled this step's buffer): signal, wait for the peers, read every rank's lists over peer memory, merge.
        Returns (scores, ids, tfidf) — or, with `deferred`, a PendingSearch of them."""
        ct = self.ctypes
        self.step += 1
        s = self.step
        base = (s % self.NBUF) * self.stride
        ps = self._arr(*[a + base + self.off_s for a in self.ptrs])
        pi = self._arr(*[a + base + self.off_i for a in self.ptrs])
        pt = self._arr(*[a + base + self.off_t for a in self.ptrs]) if with_tfidf else None
        dev = self.buf.device
        cur = torch.cuda.current_stream(dev)
        # synchronous calls stay on the caller's stream (no cross-stream hop on the latency path); merges must run in
        # step order on every rank, so such a call first waits for a still pending deferred merge
        run_on = self.merge_stream if deferred else cur
        if deferred:
            scanned = torch.cuda.Event()
            scanned.record(cur)
        else:
            prev = self._merged.get(s - 1)
            if prev is not None:
                cur.wait_event(prev)
        with torch.cuda.stream(run_on):
            if deferred:
                self.merge_stream.wait_event(scanned)
            out_s = torch.empty(B, self.k, dtype=torch.float32, device=dev)
            out_i = torch.empty(B, self.k, dtype=torch.int64, device=dev)
            out_t = torch.empty(B, self.k, dtype=torch.float64, device=dev) if with_tfidf else None
            _lib.call("ttr_topk_exchange_merge", ct.addressof(ps), ct.addressof(pi), ct.addressof(pt) if pt else None,
                      ct.addressof(self._flags), self.world, self.rank, s & 0xFFFFFFFF, B, self.k, self.k,
                      out_s, out_i, out_t)
            done = torch.cuda.Event()
            done.record(run_on)
        self._merged[s] = done
        self._merged.pop(s - 3, None)
        if deferred:
            return PendingSearch((out_s, out_i, out_t), done)
        return out_s, out_i, out_t


class ShardedIndex:
    """Row-sharded exact index: each rank keeps `docs_local` fp32 [n_local, 256] resident in HBM.

    `search` = local fused score+top-k -> all-gather of [B,k] (score, global id) over
    NCCL/NVLink -> merge kernel; the result is identical on every rank and independent of
    the number of shards.  With `group=None` and world size 1 no collective is issued."""

    def __init__(self, docs_local: torch.Tensor, row_offset: int = 0, n_total: Optional[int] = None,
                 group=None, tfidf_local: Optional[CsrF64] = None, peer_memory: bool = True):
        _lib.require_cuda(docs_local, "ShardedIndex(docs_local)")
        self.docs = docs_local.contiguous()
        self.row_offset = int(row_offset)
        self.n_total = int(n_total if n_total is not None else docs_local.shape[0])
        self.group = group
        self.tfidf = tfidf_local
        import torch.distributed as dist
        self._dist = dist
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        self.peer_memory = bool(peer_memory) and self.world > 1
        self._px: Optional[_PeerExchange] = None

    def _exchange(self, B: int, k: int) -> Optional[_PeerExchange]:
        """Symmetric-memory exchange state, created collectively on first use / growth.  The outcome is AGREED
        across the group (all-reduce MIN of an ok flag): if any rank cannot map its peers (no P2P, out of
        memory) every rank switches to the NCCL all-gather exchange — a split decision would deadlock.  B and k
        must be the same on every rank (replicated query batch); that is checked in the same exchange."""
        if not self.peer_memory:
            return None
        if self._px is None or self._px.k != k or self._px.max_b < B:
            dist = self._dist
            dev = self.docs.device
            if self._px is not None:               # growth: nothing of the old exchange may still be in flight here
                self._px.merge_stream.synchronize()
                torch.cuda.current_stream(dev).synchronize()
            ok, err = 1, None
            try:
                px = _PeerExchange(self.group, dev, max(B, 128), k)
            except Exception as e:     # no P2P mapping on this box
                ok, err, px = 0, e, None
            agree = torch.tensor([ok, B, -B, k, -k], dtype=torch.int64, device=dev)
            dist.all_reduce(agree, op=dist.ReduceOp.MIN, group=self.group)
            a = agree.tolist()
            if a[1] != -a[2] or a[3] != -a[4]:
                raise _lib.TTRError(f"ShardedIndex.search: query batch / k differ across ranks (min B {a[1]}, max B {-a[2]}, "
                                    f"min k {a[3]}, max k {-a[4]}); every rank must pass the same replicated batch")
            if a[0] == 0:
                import warnings
                warnings.warn(f"peer-memory exchange unavailable on at least one rank ({err!r} here); "
                              "all ranks use the NCCL all-gather exchange")
                self.peer_memory = False
                self._px = None
                return None
            self._px = px
        return self._px

    def _local(self, Q, k, out=None):
        if self.docs.shape[0] == 0:
            B = Q.shape[0]
            s = torch.full((B, k), float("-inf"), device=Q.device)
            i = torch.full((B, k), -1, dtype=torch.int64, device=Q.device)
            if out is not None:
                out[0].copy_(s)
                out[1].copy_(i)
                return out
            return s, i
        return search_topk(Q, self.docs, k, self.row_offset, out=out)

    def search_deferred(self, Q: torch.Tensor, k: int = 50) -> PendingSearch:
        """`search` for a stream of independent query batches: the local scan is enqueued on the caller's stream,
        the cross-rank exchange + merge on a side stream, and the call returns a PendingSearch whose `result()`
        gives (scores, ids).  The caller may enqueue the next batch right away; a rank whose peers are still
        scanning runs one scan ahead instead of waiting in the merge.  Same results as `search`."""
        if Q.dim() == 1:
            Q = Q.unsqueeze(0)
        B = Q.shape[0]
        px = self._exchange(B, k) if self.world > 1 else None
        if px is None:
            return PendingSearch(self.search(Q, k))
        vs, vi, _ = px.views(B)
        self._local(Q, k, out=(vs, vi))
        pend = px.merge(B, with_tfidf=False, deferred=True)
        return PendingSearch(pend._tensors[:2], pend._event)

    def search(self, Q: torch.Tensor, k: int = 50):
        if self.world == 1:
            return self._local(Q, k)
        if Q.dim() == 1:
            Q = Q.unsqueeze(0)
        B = Q.shape[0]
        px = self._exchange(B, k)
        if px is not None:
            vs, vi, _ = px.views(B)
            self._local(Q, k, out=(vs, vi))
            ms, mi, _ = px.merge(B, with_tfidf=False)
            return ms, mi
        s, i = self._local(Q, k)
        gs = torch.empty(self.world, B, k, dtype=s.dtype, device=s.device)
        gi = torch.empty(self.world, B, k, dtype=i.dtype, device=i.device)
        self._dist.all_gather_into_tensor(gs, s, group=self.group)
        self._dist.all_gather_into_tensor(gi, i, group=self.group)
        return topk_merge(gs, gi, k)

    def search_hybrid(self, Q: torch.Tensor, q_csr: CsrF64, alpha: float, k: int = 50, top_n: int = 10,
                      space="l2", q_sqnorm: Optional[torch.Tensor] = None):
        """Dense top-k -> TF-IDF of the candidates (each rank scores the rows it owns) ->
        blend -> top_n.  Candidate TF-IDF scores travel with the candidates in the same
        all-gather, so no rank needs another rank's CSR slice."""
        if self.tfidf is None:
            raise ValueError("ShardedIndex was built without a TF-IDF shard")
        if Q.dim() == 1:
            Q = Q.unsqueeze(0)
        px = self._exchange(Q.shape[0], k) if self.world > 1 else None
        if px is not None:
            B = Q.shape[0]
            vs, vi, vt = px.views(B)
            self._local(Q, k, out=(vs, vi))
            tfidf_candidates(vi, self.tfidf, q_csr, out=vt)
            s, i, tf = px.merge(B, with_tfidf=True)
            out = hybrid_rerank(i, s, alpha, tfidf=tf, space=space, top_n=top_n, q_sqnorm=q_sqnorm)
            out["dense_scores"], out["dense_idx"] = s, i
            return out
        s, i = self._local(Q, k)
        tf = tfidf_candidates(i, self.tfidf, q_csr)
        if self.world > 1:
            B = s.shape[0]
            gs = torch.empty(self.world, B, k, dtype=s.dtype, device=s.device)
            gi = torch.empty(self.world, B, k, dtype=i.dtype, device=i.device)
            gt = torch.empty(self.world, B, k, dtype=tf.dtype, device=tf.device)
            self._dist.all_gather_into_tensor(gs, s, group=self.group)
            self._dist.all_gather_into_tensor(gi, i, group=self.group)
            self._dist.all_gather_into_tensor(gt, tf, group=self.group)
            # merge on (score, source position): shards are ascending row ranges and each list is
            # sorted by (score desc, id asc), so source order == id order among equal scores
            src = torch.arange(self.world * k, device=s.device, dtype=torch.int64).view(self.world, 1, k)
            src = src.expand(self.world, B, k).contiguous()
            ms, msrc = topk_merge(gs, src, k)
            flat_i = gi.permute(1, 0, 2).reshape(B, self.world * k)
            flat_t = gt.permute(1, 0, 2).reshape(B, self.world * k)
            valid = msrc >= 0
            msrc_c = msrc.clamp_min(0)
            i = torch.where(valid, torch.gather(flat_i, 1, msrc_c), torch.full_like(msrc, -1))
            tf = torch.where(valid, torch.gather(flat_t, 1, msrc_c), torch.zeros_like(ms, dtype=torch.float64))
            s = ms
        out = hybrid_rerank(i, s, alpha, tfidf=tf, space=space, top_n=top_n, q_sqnorm=q_sqnorm)
        out["dense_scores"], out["dense_idx"] = s, i
        return out
