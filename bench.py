#!/usr/bin/env python3
"""bench.py — headline benchmark of the hot path (BASELINE.json `metric`):

    queries/sec, exact cosine top-50 over 8,841,823 x 256 fp32 document embeddings,
    row-sharded over N B200s (+ % of the measured HBM roofline); doc-encode passages/s.

One "step" = one pass of the search path over one query batch:
  value : whole-job queries/s with the query batch already resident in HBM
  e2e   : same through the public API with HOST (pinned) queries -> H2D -> search -> D2H
The document matrix is the resident index (the reference keeps its Chroma collection in
memory the same way, frontend/main.py:62-77); it is 9.05 GB, i.e. every pass streams far
more than the 126 MB L2, so no explicit L2 flush is needed between iterations.

After the timed loops the results are VERIFIED at every N (a handful of queries against a chunked
fp64 torch reference over all shards, and bit-identity of the result on every rank): the line carries
`"verified": true`, a mismatch exits non-zero.

`legs` holds one entry per BASELINE.json config, each with its own roofline and (N = 1) CPU baseline
and torch-on-CUDA bar:
  config2_*   1 M docs, query batch 1 and 256, one GPU
  config3_*   bulk doc encode, 8,841,823 / 8 passages per GPU from HOST token rows through `encode_rows`
  config4_*   4096 queries x 8.84 M docs, top-50 + TF-IDF hybrid rerank (`ShardedIndex.search_hybrid`)
  config5_*   data-parallel triplet step, 1024 triplets per GPU, all-reduce + fused clip/Adam
  search_b*   the headline corpus at the other BASELINE batches (1, 256, 4096)

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl reference]
N > 1 is launched by torchrun (one rank per GPU, NCCL).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

N_DOCS = 8_841_823          # MS MARCO passage count
DIM = 256
TOPK = 50
BYTES_PER_DOC = DIM * 4     # 1,024 B per document per query-batch pass
DTYPE = "f32 storage / tf32 MMA (B > 4), f32 FMA (B <= 4)"
METRIC = "queries/sec exact cosine top-50 over 8.8M docs"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=128, help="queries per step")
    ap.add_argument("--docs", type=int, default=N_DOCS)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-extra", action="store_true", help="skip the per-config legs")
    ap.add_argument("--legs", default="all", help="comma list of leg groups: search,config2,config3,config4,config5,bars")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--encode-passages", type=int, default=0, help="passages per GPU in the config-3 leg (0: 8,841,823 / 8)")
    ap.add_argument("--debug-flags", type=int, default=0, help="ttr_debug_set_flags value (tuning experiments only)")
    return ap.parse_args()


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return d, "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1900.0, "bf16_tflops_sustained": 1500.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled every 50 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: [self.lines.append(l) for l in self.proc.stdout], daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return None
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def make_shard(n_rows: int, seed: int, device) -> torch.Tensor:
    """F.normalize(N(0,1)) rows generated on the device in chunks (seed 3 + rank)."""
    out = torch.empty(n_rows, DIM, dtype=torch.float32, device=device)
    gen = torch.Generator(device=device).manual_seed(seed)
    step = 1 << 20
    for lo in range(0, n_rows, step):
        hi = min(n_rows, lo + step)
        x = torch.randn(hi - lo, DIM, device=device, generator=gen)
        out[lo:hi] = torch.nn.functional.normalize(x, dim=1)
    return out


def make_queries(batch: int, n_sets: int) -> torch.Tensor:
    gen = torch.Generator().manual_seed(4)
    q = torch.nn.functional.normalize(torch.randn(n_sets, batch, DIM, generator=gen), dim=2)
    return q


def make_csr_device(n_rows: int, device, n_features: int = 20000, mean_nnz: float = 30.0, seed: int = 5,
                    chunk_rows: int = 1 << 20):
    """Synthetic L2-normalised TF-IDF CSR generated ON THE DEVICE (F = 20,000, nnz/doc ~ Poisson(30),
    values normalised per row): indptr int64, indices int32 strictly increasing per row, data fp64 — the layout
    sklearn's TfidfVectorizer produces (backend/main.py:142-143).  Bench input only."""
    gen = torch.Generator(device=device).manual_seed(seed)
    nnz = torch.poisson(torch.full((n_rows,), float(mean_nnz), device=device), generator=gen).long().clamp_(1, n_features)
    indptr = torch.zeros(n_rows + 1, dtype=torch.int64, device=device)
    torch.cumsum(nnz, 0, out=indptr[1:])
    total = int(indptr[-1])
    indices = torch.empty(total, dtype=torch.int32, device=device)
    data = torch.empty(total, dtype=torch.float64, device=device)
    for r0 in range(0, n_rows, chunk_rows):
        r1 = min(n_rows, r0 + chunk_rows)
        lo, hi = int(indptr[r0]), int(indptr[r1])
        k = nnz[r0:r1]
        seg = torch.repeat_interleave(torch.arange(r1 - r0, device=device), k)
        pos = torch.arange(hi - lo, device=device) - (indptr[r0:r1] - lo)[seg]
        # k distinct sorted columns per row: scaled cumulative gaps in [0, F - k) plus the position
        gaps = torch.rand(hi - lo, device=device, dtype=torch.float64, generator=gen) + 0.05
        cs = torch.cumsum(gaps, 0)
        row_start = torch.zeros(r1 - r0, dtype=torch.float64, device=device)
        row_start[1:] = cs[(indptr[r0 + 1:r1] - lo - 1)]
        row_tot = torch.zeros(r1 - r0, dtype=torch.float64, device=device).index_add_(0, seg, gaps)
        frac = (cs - gaps - row_start[seg]) / row_tot[seg]
        col = (frac * (n_features - k[seg]).double()).floor().long().clamp_(min=0) + pos
        indices[lo:hi] = col.clamp_(max=n_features - 1).int()
        v = torch.rand(hi - lo, device=device, dtype=torch.float64, generator=gen) + 0.05
        nrm = torch.zeros(r1 - r0, dtype=torch.float64, device=device).index_add_(0, seg, v * v).sqrt_()
        data[lo:hi] = v / nrm[seg]
    return indptr, indices, data


def make_query_csr(batch: int, device, n_features: int = 20000, seed: int = 6):
    """Query TF-IDF rows: 3-6 distinct features each, L2-normalised."""
    rng = np.random.default_rng(seed)
    nnz = rng.integers(3, 7, size=batch)
    indptr = np.zeros(batch + 1, dtype=np.int64)
    np.cumsum(nnz, out=indptr[1:])
    idx = np.concatenate([np.sort(rng.choice(n_features, size=int(k), replace=False)) for k in nnz]).astype(np.int32)
    val = rng.random(int(indptr[-1])) + 0.05
    for b in range(batch):
        val[indptr[b]:indptr[b + 1]] /= np.linalg.norm(val[indptr[b]:indptr[b + 1]])
    from twotowermlretrieval_b200.index import CsrF64
    return CsrF64.from_arrays(indptr, idx, val, device)


# --------------------------------------------------------------------------- reference arm / CPU baselines
def cpu_docs(n_rows: int) -> torch.Tensor:
    gen = torch.Generator().manual_seed(3)
    out = torch.empty(n_rows, DIM)
    for lo in range(0, n_rows, 1 << 20):
        hi = min(n_rows, lo + (1 << 20))
        out[lo:hi] = torch.nn.functional.normalize(torch.randn(hi - lo, DIM, generator=gen), dim=1)
    return out


def cpu_reference_qps(batch: int, n_docs_total: int, budget_s: float = 15.0, sample_docs: int = 1_000_000):
    """The reference's own CPU implementation of the path: `torch.matmul(q, D.t())` +
    `torch.topk(sim, 50)` (backend/evaluators.py:185-186) via oracle.torch_path, all host threads,
    on a bounded document sample; time scales linearly in N, so queries/s over the full corpus
    = measured / (N / sample).  (`--impl reference` runs the same call over the FULL corpus.)"""
    from oracle import torch_path
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    ns = min(sample_docs, n_docs_total)
    D = cpu_docs(ns)
    Q = make_queries(batch, 1)[0]
    torch_path.cosine_topk(Q, D, TOPK)                   # warm-up
    times, t_start = [], time.perf_counter()
    while len(times) < 3 or (time.perf_counter() - t_start < budget_s and len(times) < 200):
        t0 = time.perf_counter()
        torch_path.cosine_topk(Q, D, TOPK)
        times.append(time.perf_counter() - t0)
    best = min(times)
    qps_full = batch / best * ns / n_docs_total
    return {"value": qps_full, "unit": "queries/s", "cores": threads, "kind": "port",
            "sample": f"{batch} queries x {ns} docs (first {ns} rows of the synthetic corpus), best of {len(times)} "
                      f"= {best * 1e3:.1f} ms; scaled x{ns / n_docs_total:.4f} to {n_docs_total} docs",
            "torch_threads": torch.get_num_threads()}


def cpu_reference_encode(n_passages: int = 4096, vocab_size: int = 50000):
    """The reference's CPU doc-tower path (`backend/main.py:125-133`: batches of 64 through
    `encode_document`) via oracle.torch_path on a bounded sample, all host threads -> passages/s.
    The embedding table is cut to `vocab_size` rows for the sample (lookup cost does not depend on it)."""
    from oracle import torch_path
    from twotowermlretrieval_b200 import synth
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    cfg = synth.default_config(vocab_size=vocab_size, embed_dim=200)
    sd = torch_path.to_torch_state(synth.make_state_dict(cfg, seed=0, table_seed=1))
    ids, lens = synth.make_tokens(n_passages, "passage", vocab_size, seed=2)
    rows = [ids[i, :lens[i]].tolist() for i in range(n_passages)]
    torch_path.bulk_encode_documents(sd, cfg, rows[:64], batch_size=64)          # warm-up
    t0 = time.perf_counter()
    torch_path.bulk_encode_documents(sd, cfg, rows, batch_size=64)
    dt = time.perf_counter() - t0
    return {"value": n_passages / dt, "unit": "passages/s", "cores": threads, "kind": "port",
            "sample": f"{n_passages} synthetic passages (mean length {float(lens.mean()):.1f}), reference loop of 64 per "
                      f"batch, {dt:.2f} s"}


def cpu_reference_train(batch: int = 64, vocab_size: int = 50000):
    """One live training step of the reference on the host (`backend/main.py:244-259`, BATCH_SIZE 64) -> triplets/s."""
    from oracle import torch_path
    from twotowermlretrieval_b200 import synth
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    cfg = synth.default_config(vocab_size=vocab_size, embed_dim=200)
    cfg["DROPOUT"] = 0.0
    sd = torch_path.to_torch_state(synth.make_state_dict(cfg, seed=0, table_seed=1), requires_grad=True)
    q, _ = synth.make_tokens(batch, "query", vocab_size, seed=21)
    p, _ = synth.make_tokens(batch, "passage", vocab_size, seed=22)
    n, _ = synth.make_tokens(batch, "passage", vocab_size, seed=23)
    t = [torch.tensor(a) for a in (q, p, n)]
    st = {}
    torch_path.train_step(sd, st, cfg, *t)
    t0 = time.perf_counter()
    torch_path.train_step(sd, st, cfg, *t)
    dt = time.perf_counter() - t0
    return {"value": batch / dt, "unit": "triplets/s", "cores": threads, "kind": "port",
            "sample": f"one step at the reference's BATCH_SIZE = {batch} (dropout 0), {dt:.2f} s"}


def run_reference(args):
    """`--impl reference`: the reference's CPU path (`torch.matmul` + `torch.topk`, evaluators.py:185-186) over the
    FULL corpus on the host, every step one query batch — no extrapolation."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import torch_path
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    n_docs = args.docs
    D = cpu_docs(n_docs)
    n_sets = 4
    Qs = make_queries(args.batch, n_sets)
    for w in range(args.warmup):
        torch_path.cosine_topk(Qs[w % n_sets], D, TOPK)
    t0 = time.perf_counter()
    for s in range(args.steps):
        torch_path.cosine_topk(Qs[s % n_sets], D, TOPK)
    dt = (time.perf_counter() - t0) / max(args.steps, 1)
    qps = args.batch / dt
    sample = (f"each step = {args.batch} queries x all {n_docs} docs on the host (torch.matmul + torch.topk, "
              f"{threads} threads), measured on the full corpus")
    line = {"impl": "reference", "metric": METRIC, "value": qps,
            "unit": "queries/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": bench_config(args, args.gpus),
            "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def committed_traffic(n_docs: int, batch: int, world: int) -> dict:
    """`roofline.traffic` is NOT measured in this run (ncu cannot wrap a timed run).  For the default shape on one GPU it
    is read from the committed `ncu --set full` capture of the same kernel at the same shape and labelled as such;
    otherwise null."""
    note = {"traffic": None, "traffic_note": "not measured in this run; committed ncu captures: profiles/r2_prof_scorer_*_ncu_raw.csv"}
    cap = {(8841823, 128): "r2_prof_scorer_b128_8p8M_ncu_raw.csv", (1105228, 128): "r2_prof_scorer_b128_shard8_ncu_raw.csv"}.get((n_docs, batch))
    if world != 1 or cap is None:
        return note
    try:
        import csv
        rows = list(csv.reader(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", cap))))
        hdr, unit, val = rows[0], rows[1], rows[2]
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        tot = 0.0
        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = hdr.index(m)
            tot += float(val[i].replace(",", "")) * scale[unit[i]]
        return {"traffic": tot, "traffic_note": f"FROM THE COMMITTED CAPTURE profiles/{cap} (ncu --set full of this kernel at this "
                                                "shape, dram__bytes_read + dram__bytes_write per launch) — not measured in this run"}
    except Exception:
        return note


def bench_config(args, world):
    return {"workload": f"exact cosine top-{TOPK} over {args.docs} x {DIM} fp32 synthetic doc embeddings "
                        f"(MS MARCO passage scale), query batch {args.batch}, row-sharded over {world} GPU(s)",
            "n_docs": args.docs, "dim": DIM, "k": TOPK, "query_batch": args.batch,
            "parallelism": f"row-shard x{world}, peer-memory exchange + merge",
            "l2": "no flush: each pass streams the whole shard (>= 1.1 GB) >> 126 MB L2"}


def search_launches(B: int, shard_rows: int, world: int) -> int:
    """Kernels of one `ShardedIndex.search` call (what `ttr_score_topk` + the exchange launch)."""
    if B <= 4:
        n = 2                                   # streaming scan, merge
    elif shard_rows < 4_000_000 and B <= 384:
        n = 3                                   # init, fused sample+main scan (cooperative), select-merge
    elif B <= 512:
        n = 6                                   # init, sample scan, select-merge, seed, main scan, select-merge
    else:
        n = 3                                   # init, main scan, select-merge (the sample is no bound worth a pass)
    return n + (1 if world > 1 else 0)          # + exchange-merge


# --------------------------------------------------------------------------- B200 arm
class Ctx:
    pass


def run_b200(args):
    import torch.distributed as dist
    from twotowermlretrieval_b200.index import ShardedIndex, shard_bounds

    c = Ctx()
    c.args = args
    c.world = world = int(os.environ.get("WORLD_SIZE", "1"))
    c.rank = rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    c.dev = dev = torch.device("cuda", local)
    c.dist = dist
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    if args.debug_flags:
        from twotowermlretrieval_b200 import _lib
        _lib.call_nostream("ttr_debug_set_flags", args.debug_flags)
    c.peaks, c.peak_src = peaks()
    c.hbm_peak = float(c.peaks["hbm_gbs"])
    lo, hi = shard_bounds(args.docs, world, rank)
    c.lo, c.hi = lo, hi
    c.docs = make_shard(hi - lo, 3 + rank, dev)
    c.index = index = ShardedIndex(c.docs, lo, args.docs)
    n_sets = 8
    Qh = make_queries(args.batch, n_sets).pin_memory()
    Qd = Qh.to(dev)
    B, K, W = args.batch, args.steps, args.warmup

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, finish=None):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in range(steps):
            fn(s)
        if finish is not None:
            finish()                                # (deferred merges: the timed stream waits for the last one)
        e1.record()
        sync_all()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / steps

    c.timed, c.sync_all = timed, sync_all
    out_h = (torch.empty(B, TOPK, dtype=torch.float32).pin_memory(), torch.empty(B, TOPK, dtype=torch.int64).pin_memory())

    def step_resident(s):
        index.search(Qd[s % n_sets], TOPK)

    def step_e2e(s):
        q = Qh[s % n_sets].to(dev, non_blocking=True)
        sc, ix = index.search(q, TOPK)
        out_h[0].copy_(sc, non_blocking=True)
        out_h[1].copy_(ix, non_blocking=True)
        torch.cuda.current_stream().synchronize()      # the caller reads the result every step

    def step_local_kernel(s):
        index._local(Qd[s % n_sets], TOPK)

    # throughput mode for a stream of independent batches (N > 1): the cross-rank merge of step i runs on a side stream
    # while this stream already scans step i + 1; every result is complete when the timed region ends
    pending = []

    def step_deferred(s):
        pending.append(index.search_deferred(Qd[s % n_sets], TOPK))
        if len(pending) > 2:                       # the consumer reads results two steps behind (the scan of step s has to
            pending.pop(0).result()                # wait for the merge of step s - 2 anyway); nothing accumulates

    def finish_deferred():
        for p in pending:
            p.result()
        pending.clear()

    for s in range(max(W, 3)):
        step_resident(s)
        step_e2e(s)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_sync = timed(step_resident, K)                   # every step waits for its own cross-rank merge
    if world > 1:
        for s in range(3):
            step_deferred(s)
        finish_deferred()
        ms_def = timed(step_deferred, K, finish_deferred)
        ms_res, exchange_mode = (ms_def, "deferred (ShardedIndex.search_deferred)") if ms_def < ms_sync else (ms_sync, "synchronous")
    else:
        ms_res, exchange_mode = ms_sync, "none (one shard)"   # one GPU: no exchange, the two modes are the same launches
    ms_kern = timed(step_local_kernel, K)               # scoring + local merge kernels of one shard
    ms_e2e = timed(step_e2e, K)
    # sustained: ~1.5 s of back-to-back steps (the K-step region above is a 30 ms burst)
    n_sus = int(max(50, min(4000, 1500.0 / max(ms_res, 1e-3))))
    ms_sus_sync = timed(step_resident, n_sus)
    ms_sus_def = timed(step_deferred, n_sus, finish_deferred) if world > 1 else None
    ms_sus = min(ms_sus_sync, ms_sus_def) if world > 1 else ms_sus_sync
    clocks = sampler.stop() if rank == 0 else None

    if world == 1:
        ms_kern = min(ms_kern, ms_res)      # one GPU: the resident step IS the local call (same launches)
    shard_rows = hi - lo
    algo_bytes = shard_rows * BYTES_PER_DOC     # one HBM pass per call (query tiles share document tiles through L2)
    achieved = algo_bytes / (ms_kern * 1e-3) / 1e9
    verified = verify(c, Qd[0])
    line = {
        "metric": METRIC, "value": B / (ms_res * 1e-3), "unit": "queries/s",
        "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_res, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
        "config": bench_config(args, world),
        "e2e": {"value": B / (ms_e2e * 1e-3), "unit": "queries/s", "h2d_bytes_per_step": B * DIM * 4,
                "d2h_bytes_per_step": B * TOPK * 12, "ms_per_step": ms_e2e},
        "gpu_launches": K * search_launches(B, shard_rows, world),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": c.hbm_peak, "unit": "GB/s",
                     "frac": achieved / c.hbm_peak, **committed_traffic(args.docs, B, world),
                     "peak_source": c.peak_src,
                     "kernel": ("score_topk_stream_kernel" if B <= 4 else "score_topk_mma_kernel") +
                               " (+ topk_select_merge_kernel, ~1% of the call)",
                     "algorithmic_bytes_per_call": algo_bytes, "ms_per_call": ms_kern,
                     "passes_over_shard_per_call": 1},
        "sustained": {"ms_per_step": ms_sus, "steps": n_sus, "value": B / (ms_sus * 1e-3), "ms_per_step_synchronous": ms_sus_sync,
                      "ms_per_step_deferred": ms_sus_def},
        "exchange_mode_of_value": exchange_mode,
        "synchronous": {"ms_per_step": ms_sync, "value": B / (ms_sync * 1e-3),
                        "note": "every step waits for its own cross-rank exchange + merge before the next scan is enqueued; "
                                "`value` lets the merge of step i run behind the scan of step i + 1 (ShardedIndex.search_deferred)"},
        "verified": verified["ok"], "verification": verified,
        "clocks": clocks,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_reference_qps(B, args.docs)
    if not args.no_extra:
        line["legs"] = legs(c)
    if rank == 0:
        print(json.dumps(line))
    ok = verified["ok"] and all(v.get("verified", True) for v in line.get("legs", {}).values() if isinstance(v, dict))
    if world > 1:
        dist.destroy_process_group()
    if not ok:
        raise SystemExit("bench.py: result verification FAILED (see the 'verification' fields of the JSON line)")


# --------------------------------------------------------------------------- verification
def fp64_topk_all_shards(c, Q: torch.Tensor, docs: torch.Tensor, row_offset: int, k: int):
    """Chunked fp64 `matmul` + `topk` over this rank's rows, then (N > 1) an all-gather of the per-rank lists and a
    second top-k: the reference's `evaluators.py:185-186` in fp64 over ALL shards, ties broken by lower index."""
    Qd = Q.double()
    best_s = torch.empty(Q.shape[0], 0, dtype=torch.float64, device=Q.device)
    best_i = torch.empty(Q.shape[0], 0, dtype=torch.int64, device=Q.device)
    step = 1 << 19
    for lo in range(0, docs.shape[0], step):
        hi = min(docs.shape[0], lo + step)
        sc = Qd @ docs[lo:hi].double().t()
        s, i = torch.topk(sc, min(k, hi - lo), dim=1)
        best_s, best_i = torch.cat([best_s, s], 1), torch.cat([best_i, i + lo + row_offset], 1)
        if best_s.shape[1] > 8 * k:
            best_s, best_i = _topk_by_score_then_index(best_s, best_i, k)
    if c.world > 1:
        best_s, best_i = _topk_by_score_then_index(best_s, best_i, k)
        gs = [torch.empty_like(best_s) for _ in range(c.world)]
        gi = [torch.empty_like(best_i) for _ in range(c.world)]
        c.dist.all_gather(gs, best_s.contiguous())
        c.dist.all_gather(gi, best_i.contiguous())
        best_s, best_i = torch.cat(gs, 1), torch.cat(gi, 1)
    return _topk_by_score_then_index(best_s, best_i, k)


def _topk_by_score_then_index(s, i, k):
    o = torch.argsort(i, dim=1, stable=True)
    s, i = torch.gather(s, 1, o), torch.gather(i, 1, o)
    o = torch.argsort(s, dim=1, descending=True, stable=True)[:, :k]
    return torch.gather(s, 1, o), torch.gather(i, 1, o)


def true_scores(c, Q, idx, docs, row_offset):
    """fp64 score of every returned (query, document id) pair: each rank scores the ids it owns, all-reduce(sum)."""
    loc = idx - row_offset
    own = (idx >= 0) & (loc >= 0) & (loc < docs.shape[0])
    rows = docs[loc.clamp(0, max(docs.shape[0] - 1, 0))].double()                 # [n, k, D]
    sc = (rows * Q.double().unsqueeze(1)).sum(-1) * own
    if c.world > 1:
        c.dist.all_reduce(sc)
    return sc


def check_against_fp64(c, Q, s, i, docs, row_offset, k, rtol=1e-3, atol=1e-4):
    """north_star: scores within 1e-3 relative (+1e-4 absolute on the cosine scale) of the fp32 reference path,
    indices identical except at score ties inside that tolerance."""
    s_ref, i_ref = fp64_topk_all_shards(c, Q, docs, row_offset, k)
    tol = rtol * s_ref.abs() + atol
    err = (s.double() - s_ref).abs()
    ok_scores = bool((err <= tol).all())
    ts = true_scores(c, Q, i, docs, row_offset)
    mism = i != i_ref
    ok_idx = bool(((ts - s_ref).abs() <= tol)[mism].all()) if bool(mism.any()) else True
    uniq = all(len(set(r)) == len(r) for r in i.tolist())
    return {"ok": ok_scores and ok_idx and uniq, "queries_checked": int(Q.shape[0]), "max_score_err": float(err.max()),
            "index_mismatches_at_ties": int(mism.sum()), "indices_unique": uniq}


def same_on_all_ranks(c, *tensors) -> bool:
    if c.world == 1:
        return True
    ok = True
    for t in tensors:
        t = t.contiguous()
        g = [torch.empty_like(t) for _ in range(c.world)]
        c.dist.all_gather(g, t)
        ok = ok and all(torch.equal(g[0], x) for x in g[1:])
    flag = torch.tensor([1 if ok else 0], device=c.dev)
    c.dist.all_reduce(flag, op=c.dist.ReduceOp.MIN)
    return bool(flag.item())


def verify(c, Q, n_check: int = 8):
    """Headline result check at this N: the first `n_check` queries of the batch against fp64 over all shards, and
    the whole [B, k] result bit-identical on every rank."""
    s, i = c.index.search(Q, TOPK)
    n = min(n_check, Q.shape[0])
    out = check_against_fp64(c, Q[:n], s[:n], i[:n], c.docs, c.lo, TOPK)
    out["identical_on_all_ranks"] = same_on_all_ranks(c, s, i)
    out["ok"] = out["ok"] and out["identical_on_all_ranks"]
    out["reference"] = "chunked fp64 torch.matmul + torch.topk over all shards (evaluators.py:185-186)"
    flag = torch.tensor([1 if out["ok"] else 0], device=c.dev)
    if c.world > 1:
        c.dist.all_reduce(flag, op=c.dist.ReduceOp.MIN)
    out["ok"] = bool(flag.item())
    return out


# --------------------------------------------------------------------------- legs
def want(c, group):
    sel = c.args.legs
    return sel == "all" or group in sel.split(",")


def tf32_peak(c):
    """cuBLAS TF32 GEMM throughput on this GPU (torch.matmul fp32, allow_tf32, 8192^3, best of 10): the tensor-pipe
    denominator for the kind::tf32 scorer (BASELINE.md asks for a measured tf32 peak)."""
    if getattr(c, "_tf32_peak", None) is None:
        old = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = True
        a = torch.randn(8192, 8192, device=c.dev)
        b = torch.randn(8192, 8192, device=c.dev)
        best = 1e9
        for it in range(12):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            a @ b
            e1.record()
            torch.cuda.synchronize()
            if it >= 2:
                best = min(best, e0.elapsed_time(e1))
        torch.backends.cuda.matmul.allow_tf32 = old
        c._tf32_peak = 2 * 8192 ** 3 / (best * 1e-3) / 1e12
        del a, b
    return c._tf32_peak


def search_leg(c, index, n_rows_local, n_total, B, steps, note=""):
    """One search configuration as a leg: resident-queries timing, roofline (HBM bytes and tf32 FLOPs), launches."""
    Q = make_queries(B, 2).to(c.dev)
    for s in range(3):
        index.search(Q[s % 2], TOPK)
    ms = c.timed(lambda s: index.search(Q[s % 2], TOPK), steps)
    ms_sync = ms
    if c.world > 1:                                    # throughput mode: merge of step i behind the scan of step i + 1
        pend = []

        def fin():
            for p in pend:
                p.result()
            pend.clear()
        def dstep(s):
            pend.append(index.search_deferred(Q[s % 2], TOPK))
            if len(pend) > 2:
                pend.pop(0).result()
        ms = min(ms, c.timed(dstep, steps, fin))
    gbs = n_rows_local * BYTES_PER_DOC / (ms * 1e-3) / 1e9
    tfl = 2.0 * B * n_rows_local * DIM / (ms * 1e-3) / 1e12
    hbm_bound = B <= 256
    peak_t = tf32_peak(c)
    roof = ({"bound": "hbm", "achieved": gbs, "peak": c.hbm_peak, "unit": "GB/s", "frac": gbs / c.hbm_peak,
             "peak_source": c.peak_src} if hbm_bound else
            {"bound": "tensor", "achieved": tfl, "peak": peak_t, "unit": "TFLOP/s", "frac": tfl / peak_t,
             "peak_source": "measured in this run: torch.matmul fp32 allow_tf32 8192^3, best of 10 (cuBLAS TF32)"})
    roof.update({"traffic": None, "algorithmic_bytes_per_call": n_rows_local * BYTES_PER_DOC,
                 "algorithmic_flops_per_call": 2.0 * B * n_rows_local * DIM, "ms_per_call": ms})
    return {"queries_per_s": B * 1e3 / ms, "ms_per_step": ms, "query_batch": B, "n_docs": n_total, "n_gpus": c.world,
            "hbm_gbs_per_gpu": gbs, "tf32_tflops_per_gpu": tfl, "roofline": roof, "ms_per_step_synchronous": ms_sync,
            "gpu_launches_per_step": search_launches(B, n_rows_local, c.world), "note": note}


def bar_search(c, docs, B, steps=3, chunk=1 << 18):
    """The vendor-library bar: the reference's `torch.matmul` + `torch.topk` (evaluators.py:185-186)
    on cuda — cuBLAS + ATen topk — over the same resident matrix, chunked over documents because [B, N] fp32 does not
    fit for B = 4096 (145 GB); chunk results merged by a second topk."""
    from oracle import torch_path
    Q = make_queries(B, 1)[0].to(c.dev)
    N = docs.shape[0]

    def one():
        if B * N * 4 <= 8 << 30:
            return torch_path.cosine_topk(Q, docs, TOPK)
        ss, ii = [], []
        for lo in range(0, N, chunk):
            s, i = torch_path.cosine_topk(Q, docs[lo:lo + chunk], TOPK)
            ss.append(s); ii.append(i + lo)
        s, i = torch.cat(ss, 1), torch.cat(ii, 1)
        s2, o = torch.topk(s, TOPK, dim=1)
        return s2, torch.gather(i, 1, o)

    one()
    ms = c.timed(lambda s: one(), steps)
    return {"kind": "torch_cuda_bar", "queries_per_s": B * 1e3 / ms, "ms_per_step": ms,
            "what": "oracle.torch_path.cosine_topk on cuda (cuBLAS sgemm + at::topk), same matrix, allow_tf32 off (torch default)"}


def legs(c):
    out = {}
    a = c.args
    full = a.docs == N_DOCS
    cpu_ok = c.rank == 0 and c.world == 1 and not a.no_cpu_baseline
    # ---- the headline corpus at the other BASELINE batches
    if want(c, "search"):
        for B, steps in ((1, 20), (256, 10), (4096, 3)):
            if B == a.batch:
                continue
            try:
                leg = search_leg(c, c.index, c.hi - c.lo, a.docs, B, steps)
                if c.world == 1 and want(c, "bars"):
                    bar = bar_search(c, c.docs, B)
                    leg["torch_cuda_bar"] = bar
                    leg["vs_bar"] = leg["queries_per_s"] / bar["queries_per_s"]
                out[f"search_b{B}"] = leg
            except Exception as e:
                out[f"search_b{B}"] = {"error": repr(e)}
        if c.world == 1 and want(c, "bars"):
            try:
                out["headline_torch_cuda_bar"] = bar_search(c, c.docs, a.batch)
            except Exception as e:
                out["headline_torch_cuda_bar"] = {"error": repr(e)}
    # ---- config 2: 1 M docs, batch 1 and 256, single B200
    if want(c, "config2") and c.world == 1 and c.docs.shape[0] >= 1_000_000:
        from twotowermlretrieval_b200.index import ShardedIndex
        d1m = c.docs[:1_000_000]
        idx1m = ShardedIndex(d1m, 0, 1_000_000)
        for B, steps in ((1, 50), (256, 30)):
            try:
                leg = search_leg(c, idx1m, 1_000_000, 1_000_000, B, steps, note="BASELINE configs[1]")
                Q = make_queries(B, 2).to(c.dev)[0]
                s, i = idx1m.search(Q, TOPK)
                n = min(8, B)
                leg["verification"] = check_against_fp64(c, Q[:n], s[:n], i[:n], d1m, 0, TOPK)
                leg["verified"] = leg["verification"]["ok"]
                if want(c, "bars"):
                    bar = bar_search(c, d1m, B, steps=10)
                    leg["torch_cuda_bar"] = bar
                    leg["vs_bar"] = leg["queries_per_s"] / bar["queries_per_s"]
                if cpu_ok:
                    leg["cpu_baseline"] = cpu_reference_qps(B, 1_000_000, budget_s=5.0)
                out[f"config2_1M_b{B}"] = leg
            except Exception as e:
                out[f"config2_1M_b{B}"] = {"error": repr(e)}
    # ---- config 4: 4096 queries x 8.84 M docs, top-50 + TF-IDF hybrid rerank
    if want(c, "config4"):
        try:
            out["config4_b4096_hybrid"] = leg_hybrid(c)
        except Exception as e:
            out["config4_b4096_hybrid"] = {"error": repr(e)}
    torch.cuda.empty_cache()
    # ---- config 3: bulk doc encode from host token rows into the resident shard
    if want(c, "config3"):
        try:
            out["config3_bulk_encode"] = leg_encode(c, cpu_ok)
        except Exception as e:
            out["config3_bulk_encode"] = {"error": repr(e)}
    torch.cuda.empty_cache()
    # ---- config 5: data-parallel training step
    if want(c, "config5"):
        try:
            out["config5_dp_train_step"] = leg_train(c, cpu_ok)
        except Exception as e:
            out["config5_dp_train_step"] = {"error": repr(e)}
    return out


def leg_hybrid(c, B: int = 4096, alpha: float = 0.5):
    """BASELINE configs[3]: `ShardedIndex.search_hybrid` = dense top-50 (tcgen05 scan) -> TF-IDF cosine of the candidates
    each rank owns -> exchange + merge -> alpha blend, stable sort, top-10 (frontend/main.py:153-198), all on device."""
    from twotowermlretrieval_b200.index import CsrF64, ShardedIndex, hybrid_rerank
    n_local = c.hi - c.lo
    indptr, indices, data = make_csr_device(n_local, c.dev, seed=5 + c.rank)
    csr = CsrF64(indptr, indices, data, n_local, c.lo)
    index = ShardedIndex(c.docs, c.lo, c.args.docs, tfidf_local=csr)
    Q = make_queries(B, 1)[0].to(c.dev)
    q_csr = make_query_csr(B, c.dev)
    out = index.search_hybrid(Q, q_csr, alpha, k=TOPK, top_n=10)
    steps = 3
    ms = c.timed(lambda s: index.search_hybrid(Q, q_csr, alpha, k=TOPK, top_n=10), steps)
    ms_dense = c.timed(lambda s: index.search(Q, TOPK), steps)
    tfl = 2.0 * B * n_local * DIM / (ms * 1e-3) / 1e12
    peak_t = tf32_peak(c)
    # verification: dense candidates of 8 queries vs fp64 over all shards; the blend recomputed with torch fp64 from the
    # returned candidates and their TF-IDF scores; identical on every rank
    n = 8
    ver = check_against_fp64(c, Q[:n], out["dense_scores"][:n], out["dense_idx"][:n], c.docs, c.lo, TOPK)
    sem = 2.0 * out["dense_scores"][:n].double() - 1.0
    # TF-IDF of the returned candidates straight from the CSR rows this rank owns, all-reduced
    from twotowermlretrieval_b200.index import tfidf_candidates
    tf = tfidf_candidates(out["dense_idx"][:n].contiguous(), csr, CsrF64(q_csr.indptr[:n + 1].contiguous(), q_csr.indices,
                                                                       q_csr.data, n, 0))
    if c.world > 1:
        c.dist.all_reduce(tf)
    fin = alpha * sem + (1.0 - alpha) * tf
    order = torch.argsort(fin, dim=1, descending=True, stable=True)[:, :10]
    ref_fin = torch.gather(fin, 1, order)
    ref_idx = torch.gather(out["dense_idx"][:n], 1, order)
    blend_ok = bool(torch.allclose(out["final"][:n], ref_fin, rtol=0, atol=1e-12)) and bool(torch.equal(out["idx"][:n], ref_idx))
    same = same_on_all_ranks(c, out["idx"], out["final"])
    ver.update({"blend_matches_fp64_recompute": blend_ok, "identical_on_all_ranks": same})
    ver["ok"] = ver["ok"] and blend_ok and same
    return {"queries_per_s": B * 1e3 / ms, "ms_per_step": ms, "ms_dense_only": ms_dense, "query_batch": B,
            "n_docs": c.args.docs, "n_gpus": c.world, "alpha": alpha, "candidates": TOPK, "top_n": 10,
            "tfidf": {"features": 20000, "mean_nnz_per_doc": 30, "local_nnz": int(indptr[-1])},
            "roofline": {"bound": "tensor", "achieved": tfl, "peak": peak_t, "unit": "TFLOP/s", "frac": tfl / peak_t,
                         "traffic": None, "algorithmic_flops_per_call": 2.0 * B * n_local * DIM, "ms_per_call": ms,
                         "peak_source": "measured in this run: torch.matmul fp32 allow_tf32 8192^3, best of 10 (cuBLAS TF32)",
                         "kernel": "score_topk_mma_kernel (kind::tf32), per GPU"},
            "gpu_launches_per_step": search_launches(B, n_local, c.world) + 2,      # + tfidf_candidates, hybrid_rerank
            "verified": ver["ok"], "verification": ver, "note": "BASELINE configs[3]"}


def leg_encode(c, cpu_ok):
    """BASELINE configs[2]: the doc tower over this rank's share of the corpus, from HOST token rows (flat ids + lengths)
    through `encode.encode_rows` (length-bucketed batches, two pinned staging buffers, async H2D) straight into the rows of
    the resident search shard.  8,841,823 / 8 passages per GPU (what each of the 8 GPUs of the config encodes)."""
    from twotowermlretrieval_b200 import TwoTowerModel, _lib, synth
    from twotowermlretrieval_b200.encode import encode_rows
    n_pass = c.args.encode_passages or -(-N_DOCS // 8)
    n_pass = min(n_pass, c.docs.shape[0])
    cfg = synth.default_config()
    torch.manual_seed(0)
    table = synth.make_state_dict(dict(cfg, HIDDEN_DIM=8, NUM_LAYERS=1, BIDIRECTIONAL=False), seed=0, table_seed=1)[
        "doc_encoder.embedding.weight"]
    model = TwoTowerModel(cfg, table)
    model.to(c.dev).eval()
    enc = model.doc_encoder
    t0 = time.perf_counter()
    flat, lens = synth.make_ragged_tokens(n_pass, "passage", cfg["VOCAB_SIZE"], seed=2 + c.rank)
    gen_s = time.perf_counter() - t0
    toks = int(lens.sum())
    shard = c.docs                                      # encode into the first n_pass rows of the resident shard
    encode_rows(enc, (flat[: int(lens[:4096].sum())], lens[:4096]), c.dev, out=shard, out_offset=0)      # warm-up
    c.sync_all()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    encode_rows(enc, (flat, lens), c.dev, out=shard, out_offset=0)
    e1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    ms = torch.tensor([e0.elapsed_time(e1), wall * 1e3], device=c.dev)
    if c.world > 1:
        c.dist.all_reduce(ms, op=c.dist.ReduceOp.MAX)
    ms_dev, ms_wall = float(ms[0]), float(ms[1])
    # verification: unit-norm rows, and a 64-row sample against a batch-of-one-length encode of the same rows
    rows = torch.randint(0, n_pass, (64,), generator=torch.Generator().manual_seed(1)).tolist()
    starts = np.concatenate([[0], np.cumsum(lens)])
    sub_flat = np.concatenate([flat[starts[r]:starts[r] + lens[r]] for r in rows])
    again = encode_rows(enc, (sub_flat, lens[rows]), c.dev)
    got = shard[rows]
    norm_ok = bool(((shard[:n_pass].norm(dim=1) - 1.0).abs() < 1e-4).all())
    same_ok = bool(((got - again).norm(dim=1) <= 1e-3).all())
    # compute-only: device-resident, pre-sorted 7,680-row batches (30 cluster tiles x 2 directions x 4 waves)
    from twotowermlretrieval_b200.encode import encode_padded_batches
    NP, BS = 61440, 7680
    ids, l2 = synth.make_tokens(NP, "passage", cfg["VOCAB_SIZE"], seed=2 + c.rank)
    order = np.argsort(-l2, kind="stable")
    batches = [torch.tensor(ids[order[i:i + BS], :int(l2[order[i]])], device=c.dev) for i in range(0, NP, BS)]
    enc.strict_lengths = False
    with torch.no_grad():
        for b in batches:
            enc(b)
        encode_padded_batches(enc, batches, streams=2)
        ms_c = c.timed(lambda s: encode_padded_batches(enc, batches), 3)              # one batch at a time (default)
        ms_c2 = c.timed(lambda s: encode_padded_batches(enc, batches, streams=2), 3)  # A/B: two compute streams (slower)
    toks_c = int(l2.sum())
    # the projection GEMMs alone, same token count, own kernel time (north_star: >= 50 % tensor-pipe utilisation)
    proj = {}
    peak_bf16 = float(c.peaks.get("bf16_tflops", 1667.1))
    M = toks_c
    for name, Kd in (("l0_K200", 200), ("l1_K512", 512)):
        X = (torch.randn(M, Kd, device=c.dev) * 0.3).half()
        Wt = (torch.randn(1536, Kd, device=c.dev) * 0.06).half()
        bias = torch.zeros(1536, device=c.dev)
        gi = torch.empty(M, 1536, dtype=torch.float16, device=c.dev)
        mv = torch.tensor([M], dtype=torch.int32, device=c.dev)
        f = lambda s: _lib.call("ttr_gemm_f16_bias", X, Wt, bias, gi, M, mv, 1536, Kd)
        for s in range(3):
            f(s)
        msg = c.timed(f, 10)
        tf = 2.0 * M * Kd * 1536 / (msg * 1e-3) / 1e12
        proj[name] = {"ms": msg, "tflops": tf, "frac_of_bf16_peak": tf / peak_bf16}
        del X, Wt, gi
    fl = sum(2.0 * M * Kd * 1536 for Kd in (200, 512))
    tsum = sum(p["ms"] for p in proj.values())
    proj_tf = fl / (tsum * 1e-3) / 1e12
    leg = {"passages_per_s": n_pass * c.world / (ms_wall * 1e-3), "passages_per_s_device_timed": n_pass * c.world / (ms_dev * 1e-3),
           "passages_per_gpu": n_pass, "tokens_per_gpu": toks, "mean_len": float(lens.mean()), "n_gpus": c.world,
           "ms_wall": ms_wall, "ms_device": ms_dev, "host_token_generation_s": gen_s,
           "h2d_bytes": toks * 8, "path": "host (flat ids, lengths) -> encode_rows -> rows of the resident search shard",
           "compute_only": {"passages_per_s": NP * c.world / (ms_c * 1e-3), "tokens_per_s": toks_c * c.world / (ms_c * 1e-3),
                            "ms_per_61440_passages": ms_c, "whole_tower_tflops_per_gpu": toks_c * 3_760_128.0 / (ms_c * 1e-3) / 1e12,
                            "two_streams": {"passages_per_s": NP * c.world / (ms_c2 * 1e-3), "ms_per_61440_passages": ms_c2},
                            "note": "device-resident ids, pre-sorted 7,680-row batches (encode.encode_padded_batches); "
                                    "two_streams = consecutive batches on alternating compute streams (A/B, not the default)"},
           "roofline": {"bound": "tensor", "achieved": proj_tf, "peak": peak_bf16, "unit": "TFLOP/s", "frac": proj_tf / peak_bf16,
                        "traffic": None, "kernel": "gemm_bias_kernel<F16> (input projections, kind::f16), own kernel time on the "
                        "token count of the compute-only batches", "per_layer": proj, "peak_source": c.peak_src + " bf16 burst"},
           "verified": norm_ok and same_ok, "verification": {"unit_norm_rows": norm_ok, "sample_rows_match_re_encode": same_ok},
           "config": "GRU 2-layer bidirectional H=256 E=200 V=400005 (backend/config.json); fp16 storage of X / gi / inter-layer y, "
                     "fp32 accumulation and state", "note": "BASELINE configs[2]: each of 8 GPUs encodes 8,841,823/8 passages"}
    if cpu_ok:
        try:
            leg["cpu_baseline"] = cpu_reference_encode()
        except Exception as e:
            leg["cpu_baseline"] = {"error": repr(e)}
    if c.world == 1 and want(c, "bars"):
        try:
            leg["torch_cuda_bar"] = bar_encode(c, cfg, batches, NP)
            leg["vs_bar"] = leg["compute_only"]["passages_per_s"] / leg["torch_cuda_bar"]["passages_per_s"]
        except Exception as e:
            leg["torch_cuda_bar"] = {"error": repr(e)}
    return leg


def bar_encode(c, cfg, batches, n_pass):
    """Vendor-library bar for the tower: the reference forward (`model.py:48-75`: nn.Embedding, pack_padded_sequence,
    cuDNN GRU, Linear, normalize) on cuda over the same 7,680-row batches."""
    from oracle import torch_path
    from twotowermlretrieval_b200 import synth
    sd = {k: v.to(c.dev) for k, v in torch_path.to_torch_state(synth.make_state_dict(cfg, seed=0, table_seed=1)).items()
          if k.startswith("doc_encoder")}
    with torch.no_grad():
        for b in batches:
            torch_path.encoder_forward(sd, "doc_encoder", b, cfg)
        ms = c.timed(lambda s: [torch_path.encoder_forward(sd, "doc_encoder", b, cfg) for b in batches], 3)
    return {"kind": "torch_cuda_bar", "passages_per_s": n_pass * 1e3 / ms, "ms_per_61440_passages": ms,
            "what": "oracle.torch_path.encoder_forward on cuda (cuDNN GRU, fp32), same batches"}


def leg_train(c, cpu_ok, per_gpu: int = 1024):
    """BASELINE configs[4]: one data-parallel step = 3 encodes + cosine triplet loss + backward + all-reduce of the flat
    16 MB gradient bucket + fused clip(1.0) + Adam (backend/main.py:244-259), 1024 triplets per GPU."""
    from twotowermlretrieval_b200 import TwoTowerModel, synth
    from twotowermlretrieval_b200.trainer import TrainerFactory
    cfg = synth.default_config()
    cfg["DROPOUT"] = 0.0                                   # deterministic step (train-mode dropout timing differs by one small kernel)
    sd = synth.make_state_dict(cfg, seed=0, table_seed=1)
    model = TwoTowerModel(cfg, sd["query_encoder.embedding.weight"])
    model.load_state_dict({k: torch.tensor(v) for k, v in sd.items()})
    model.to(c.dev).train()
    model.query_encoder.strict_lengths = model.doc_encoder.strict_lengths = False
    # the public step: TwoTowerTrainer.train_step = 3 encodes (three tower streams) -> loss -> backward -> all-reduce ->
    # fused clip + Adam
    trainer = TrainerFactory.create_trainer(cfg, model, c.dev, fused=True, clip_max_norm=1.0)
    V = cfg["VOCAB_SIZE"]
    t = [torch.tensor(synth.make_tokens(per_gpu, kind, V, seed=31 + j + 10 * c.rank)[0], device=c.dev)
         for j, kind in enumerate(("query", "passage", "passage"))]
    ntok = int(sum((x != 0).sum() for x in t))

    def step(s):
        return trainer.train_step(*t)[0]

    for s in range(2):
        step(s)
    ms = c.timed(step, 5)
    lanes = trainer.tower_streams
    trainer.tower_streams = 1
    step(0)
    ms_one = c.timed(step, 5)
    trainer.tower_streams = lanes
    # verification: every rank holds the same parameters after the all-reduced steps; loss is finite
    same = same_on_all_ranks(c, model.flat_params())
    finite = bool(torch.isfinite(step(0)))
    flops = ntok * 11.3e6
    leg = {"triplets_per_s": per_gpu * c.world * 1e3 / ms, "ms_per_step": ms, "triplets_per_gpu": per_gpu, "n_gpus": c.world,
           "global_batch": per_gpu * c.world, "tokens_per_gpu": ntok, "algorithmic_tflops_per_gpu": flops / (ms * 1e-3) / 1e12,
           "allreduce_bytes": int(model.flat_grads().numel() * 4), "tower_streams": lanes,
           "one_stream": {"ms_per_step": ms_one, "triplets_per_s": per_gpu * c.world * 1e3 / ms_one},
           "verified": same and finite,
           "verification": {"parameters_identical_on_all_ranks": same, "loss_finite": finite},
           "roofline": {"bound": "tensor", "achieved": flops / (ms * 1e-3) / 1e12, "peak": float(c.peaks.get("bf16_tflops", 1667.1)),
                        "unit": "TFLOP/s", "frac": flops / (ms * 1e-3) / 1e12 / float(c.peaks.get("bf16_tflops", 1667.1)),
                        "traffic": None, "note": "latency-bound recurrence chains; SURVEY 8(d) asks for step ms and triplets/s, "
                        "no roofline claim"},
           "note": "BASELINE configs[4] (DROPOUT 0)"}
    if cpu_ok:
        try:
            leg["cpu_baseline"] = cpu_reference_train()
        except Exception as e:
            leg["cpu_baseline"] = {"error": repr(e)}
    if c.world == 1 and want(c, "bars"):
        try:
            from oracle import torch_path
            sdt = {k: v.detach().to(c.dev).requires_grad_(v.requires_grad) for k, v in
                   torch_path.to_torch_state(sd, requires_grad=True).items()}
            st = {}
            for s in range(2):
                torch_path.train_step(sdt, st, cfg, *t)
            msb = c.timed(lambda s: torch_path.train_step(sdt, st, cfg, *t), 3)
            leg["torch_cuda_bar"] = {"kind": "torch_cuda_bar", "triplets_per_s": per_gpu * 1e3 / msb, "ms_per_step": msb,
                                     "what": "oracle.torch_path.train_step on cuda (cuDNN GRU fwd+bwd, torch Adam, clip_grad_norm_), "
                                             "same 1024 triplets"}
            leg["vs_bar"] = leg["triplets_per_s"] / leg["torch_cuda_bar"]["triplets_per_s"]
        except Exception as e:
            leg["torch_cuda_bar"] = {"error": repr(e)}
    return leg


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
        run_b200(args)


if __name__ == "__main__":
    main()
