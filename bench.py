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


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=128, help="queries per step")
    ap.add_argument("--docs", type=int, default=N_DOCS)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary measurements")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--debug-flags", type=int, default=0, help="ttr_debug_set_flags value (tuning experiments only)")
    return ap.parse_args()


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)", d
    return 6650.0, "fallback (B200_PROFILING.md)", {}


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


def ncu_traffic(batch: int, shard_rows: int):
    """dram__bytes_read.sum + dram__bytes_write.sum of the scoring kernel from the committed
    `ncu --set full` capture (profiles/r1_score_topk_mma_v6_ncu_raw.csv) — only valid for the exact
    configuration that was captured (B=128, full corpus on one GPU); null otherwise."""
    p = ROOT / "profiles" / "r1_score_topk_mma_v6_ncu_raw.csv"
    if batch != 128 or shard_rows != N_DOCS or not p.exists():
        return None
    try:
        import csv
        rows = list(csv.reader(open(p)))
        hdr, unit, last = rows[0], rows[1], rows[-1]
        tot = 0.0
        for name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = hdr.index(name)
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[unit[i]]
            tot += float(last[i]) * scale
        return tot
    except Exception:
        return None


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


# --------------------------------------------------------------------------- reference arm
def cpu_reference_qps(batch: int, n_docs_total: int, budget_s: float = 20.0, sample_docs: int = 1_000_000):
    """The reference's own CPU implementation of the path: `torch.matmul(q, D.t())` +
    `torch.topk(sim, 50)` (backend/evaluators.py:185-186) via oracle.torch_path, all host threads,
    on a bounded document sample; time scales linearly in N, so queries/s over the full corpus
    = measured / (N / sample)."""
    from oracle import torch_path
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    ns = min(sample_docs, n_docs_total)
    gen = torch.Generator().manual_seed(3)
    D = torch.nn.functional.normalize(torch.randn(ns, DIM, generator=gen), dim=1)
    Q = make_queries(batch, 1)[0]
    torch_path.cosine_topk(Q, D, TOPK)                   # warm-up
    times, t_start = [], time.perf_counter()
    while len(times) < 3 or (time.perf_counter() - t_start < budget_s and len(times) < 200):
        t0 = time.perf_counter()
        torch_path.cosine_topk(Q, D, TOPK)
        times.append(time.perf_counter() - t0)
    best = min(times)
    qps_sample = batch / best
    qps_full = qps_sample * ns / n_docs_total
    return {"value": qps_full, "unit": "queries/s", "cores": threads, "kind": "port",
            "sample": f"{batch} queries x {ns} docs (first {ns} rows of the synthetic corpus), best of {len(times)} "
                      f"= {best * 1e3:.1f} ms; scaled x{ns / n_docs_total:.4f} to {n_docs_total} docs",
            "torch_threads": torch.get_num_threads()}, best, ns


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


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_docs = args.docs
    from oracle import torch_path
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    ns = min(1_000_000, n_docs)
    gen = torch.Generator().manual_seed(3)
    D = torch.nn.functional.normalize(torch.randn(ns, DIM, generator=gen), dim=1)
    Qs = make_queries(args.batch, max(args.steps, 1))
    for w in range(args.warmup):
        torch_path.cosine_topk(Qs[w % len(Qs)], D, TOPK)
    t0 = time.perf_counter()
    for s in range(args.steps):
        torch_path.cosine_topk(Qs[s % len(Qs)], D, TOPK)
    dt = (time.perf_counter() - t0) / max(args.steps, 1)
    step_full = dt * n_docs / ns                    # one step over the full corpus, extrapolated linearly in N
    qps = args.batch / step_full
    sample = (f"each step = {args.batch} queries x {ns} docs on the host (torch.matmul + torch.topk, "
              f"{threads} threads), scaled x{n_docs / ns:.3f} to {n_docs} docs")
    line = {"impl": "reference", "metric": "queries/sec exact cosine top-50 over 8.8M docs", "value": qps,
            "unit": "queries/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": step_full * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": bench_config(args, 1),
            "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def bench_config(args, world):
    return {"workload": f"exact cosine top-{TOPK} over {args.docs} x {DIM} fp32 synthetic doc embeddings "
                        f"(MS MARCO passage scale), query batch {args.batch}, row-sharded over {world} GPU(s)",
            "n_docs": args.docs, "dim": DIM, "k": TOPK, "query_batch": args.batch,
            "parallelism": f"row-shard x{world} + all-gather merge" if world > 1 else "single GPU",
            "l2": "no flush: each pass streams the whole shard (>= 1.1 GB) >> 126 MB L2"}


# --------------------------------------------------------------------------- B200 arm
def run_b200(args):
    import torch.distributed as dist
    from twotowermlretrieval_b200.index import ShardedIndex, shard_bounds

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    if args.debug_flags:
        from twotowermlretrieval_b200 import _lib
        _lib.call_nostream("ttr_debug_set_flags", args.debug_flags)
    lo, hi = shard_bounds(args.docs, world, rank)
    docs = make_shard(hi - lo, 3 + rank, dev)
    index = ShardedIndex(docs, lo, args.docs)
    n_sets = 8
    Qh = make_queries(args.batch, n_sets).pin_memory()
    Qd = Qh.to(dev)
    B, K, W = args.batch, args.steps, args.warmup

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in range(steps):
            fn(s)
        e1.record()
        sync_all()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / steps

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

    for s in range(max(W, 3)):
        step_resident(s)
        step_e2e(s)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_res = timed(step_resident, K)
    ms_kern = timed(step_local_kernel, K)               # scoring + local merge kernels of one shard
    ms_e2e = timed(step_e2e, K)
    clocks = sampler.stop() if rank == 0 else None

    if world == 1:
        ms_kern = min(ms_kern, ms_res)      # one GPU: the resident step IS the local call (same launches)
    hbm_peak, peak_src, _ = peaks()
    shard_rows = hi - lo
    # B <= 4: CUDA-core streaming kernel, one pass; B > 4: tcgen05 kernel, one HBM pass per call
    # (query tiles of 128 share document tiles through L2)
    passes = 1
    algo_bytes = shard_rows * BYTES_PER_DOC * passes     # per search call on this rank
    achieved = algo_bytes / (ms_kern * 1e-3) / 1e9
    line = {
        "metric": "queries/sec exact cosine top-50 over 8.8M docs", "value": B / (ms_res * 1e-3), "unit": "queries/s",
        "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_res, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(args, world),
        "e2e": {"value": B / (ms_e2e * 1e-3), "unit": "queries/s", "h2d_bytes_per_step": B * DIM * 4,
                "d2h_bytes_per_step": B * TOPK * 12, "ms_per_step": ms_e2e},
        # B <= 4: streaming kernel + merge; 4 < B <= 128 on shards < 4 M docs: init, fused sample+main scan,
        # select-merge; otherwise: init, sample scan, merge, seed, main scan, merge; + barrier/peer merge at N > 1
        "gpu_launches": K * ((2 if B <= 4 else (3 if (B <= 128 and shard_rows < 4_000_000) else 6)) + (2 if world > 1 else 0)),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                     "frac": achieved / hbm_peak, "traffic": ncu_traffic(B, shard_rows), "peak_source": peak_src,
                     "kernel": ("score_topk_stream_kernel" if B <= 4 else "score_topk_mma_kernel") +
                               " (+ topk_merge_kernel, ~1% of the call)",
                     "algorithmic_bytes_per_call": algo_bytes, "ms_per_call": ms_kern,
                     "passes_over_shard_per_call": passes},
        "clocks": clocks,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cb, _, _ = cpu_reference_qps(B, args.docs)
        line["cpu_baseline"] = cb
    if not args.no_extra:
        line["extra"] = extras(index, dev, world, rank, timed, hbm_peak)
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            try:
                line["extra"]["doc_encode_cpu_baseline"] = cpu_reference_encode()
            except Exception as e:
                line["extra"]["doc_encode_cpu_baseline"] = {"error": repr(e)}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def extras(index, dev, world, rank, timed, hbm_peak):
    """Secondary measurements (same run, not the headline): batch-1 search latency and the
    document-tower bulk encode rate at config.json dims."""
    out = {}
    q1 = make_queries(1, 4).to(dev)
    for s in range(20):                      # the CPU baseline left the GPU idle for ~20 s: let the clocks come back
        index.search(q1[s % 4], TOPK)
    ms = timed(lambda s: index.search(q1[s % 4], TOPK), 20)
    rows = index.docs.shape[0]
    out["search_batch1"] = {"queries_per_s": 1e3 / ms, "ms": ms,
                            "hbm_gbs_per_gpu": rows * BYTES_PER_DOC / (ms * 1e-3) / 1e9,
                            "hbm_frac": rows * BYTES_PER_DOC / (ms * 1e-3) / 1e9 / hbm_peak,
                            "note": "includes the cross-rank merge at N > 1"}
    q4k = make_queries(4096, 1).to(dev)
    index.search(q4k[0], TOPK)
    ms = timed(lambda s: index.search(q4k[0], TOPK), 3)
    flops = 2.0 * 4096 * N_DOCS * DIM
    out["search_batch4096"] = {"queries_per_s": 4096e3 / ms, "ms": ms, "tf32_tflops_whole_job": flops / (ms * 1e-3) / 1e12,
                               "note": "BASELINE configs[3] shape; tensor/epilogue-bound regime (kind::tf32)"}
    try:
        from twotowermlretrieval_b200 import TwoTowerModel, synth
        cfg = synth.default_config()
        torch.manual_seed(0)
        model = TwoTowerModel(cfg, None)
        model.doc_encoder.embedding.weight.requires_grad_(False)
        model.to(dev).eval()
        model.doc_encoder.strict_lengths = False
        # 7,680 rows = 30 cluster tiles x 2 directions = 4 full waves of the 15 eight-CTA clusters a B200 holds
        NP, BS = 15360, 7680
        ids, lens = synth.make_tokens(NP, "passage", cfg["VOCAB_SIZE"], seed=2 + rank)
        order = np.argsort(-lens, kind="stable")
        batches = [torch.tensor(ids[order[i:i + BS], :int(lens[order[i]])], device=dev) for i in range(0, NP, BS)]
        with torch.no_grad():
            for b in batches:
                model.encode_document(b)
            ms = timed(lambda s: [model.encode_document(b) for b in batches], 3)
        toks = int(lens.sum())
        proj_flops = toks * 2_187_264.0          # input-projection GEMMs, both layers and directions
        out["doc_encode"] = {"passages_per_s": NP * world / (ms * 1e-3), "tokens_per_s": toks * world / (ms * 1e-3),
                             "ms_per_15360_passages": ms, "mean_len": float(lens.mean()),
                             "whole_tower_tflops": toks * 3_760_128.0 / (ms * 1e-3) / 1e12,
                             "projection_flop_share_tflops": proj_flops / (ms * 1e-3) / 1e12,
                             "config": "GRU 2-layer bidirectional H=256 E=200 V=400005 (backend/config.json), "
                                       "length-sorted batches of 7680 passages, device-resident ids; fp16 storage of X / gi / "
                                       "inter-layer y, fp32 accumulation and state; projections and recurrence on tcgen05 "
                                       "(kind::f16)"}
    except Exception as e:  # secondary measurement must never kill the headline line
        out["doc_encode"] = {"error": repr(e)}
    return out


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
