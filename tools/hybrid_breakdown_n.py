"""torchrun -n N: per-stage CUDA-event timing of the sharded `search_hybrid` at B = 4096 (r2: 55 ms at 2 GPUs against
20 ms for the dense search alone)."""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, ".")
import bench
from twotowermlretrieval_b200.index import CsrF64, ShardedIndex, hybrid_rerank, tfidf_candidates, shard_bounds
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 8_841_823
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
lo, hi = shard_bounds(N, world, rank)
docs = bench.make_shard(hi - lo, 3 + rank, dev)
ip, ix, dv = bench.make_csr_device(hi - lo, dev, seed=5 + rank)
csr = CsrF64(ip, ix, dv, hi - lo, lo)
index = ShardedIndex(docs, lo, N, tfidf_local=csr)
Q = bench.make_queries(B, 1)[0].to(dev)
qc = bench.make_query_csr(B, dev)
def t(fn, n=3):
    fn(); torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): r = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, r
ms_l, (s, i) = t(lambda: index._local(Q, 50))
ms_d, (s, i) = t(lambda: index.search(Q, 50))
px = index._exchange(B, 50)
def stage_local():
    vs, vi, vt = px.views(B); return index._local(Q, 50, out=(vs, vi))
def stage_tfidf():
    vs, vi, vt = px.views(B); return tfidf_candidates(vi, csr, qc, out=vt)
ms_a, _ = t(stage_local)
ms_t, _ = t(stage_tfidf)
ms_m, _ = t(lambda: px.merge(B, with_tfidf=False))
ms_mt, (ms_, mi_, mt_) = t(lambda: px.merge(B, with_tfidf=True))
ms_r, _ = t(lambda: hybrid_rerank(mi_, ms_, 0.5, tfidf=mt_, top_n=10))
ms_h, _ = t(lambda: index.search_hybrid(Q, qc, 0.5, k=50, top_n=10))
if rank == 0:
    print(f"world {world} N={N} B={B}: local {ms_l:.3f}, search {ms_d:.3f}, local->symm {ms_a:.3f}, tfidf->symm {ms_t:.3f}, "
          f"merge {ms_m:.3f}, merge+tfidf {ms_mt:.3f}, rerank {ms_r:.3f}, search_hybrid {ms_h:.3f} ms")
dist.destroy_process_group()
