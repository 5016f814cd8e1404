"""Per-stage CUDA-event timing of `ShardedIndex.search_hybrid` at B = 4096 (one GPU): dense scan, candidate TF-IDF,
rerank — to see which stage costs what (r2: the hybrid leg took 55 ms against 20 ms for the dense scan alone)."""
import sys, torch
sys.path.insert(0, ".")
import bench
from twotowermlretrieval_b200.index import CsrF64, ShardedIndex, hybrid_rerank, tfidf_candidates
dev = torch.device("cuda:0")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
docs = bench.make_shard(N, 3, dev)
ip, ix, dv = bench.make_csr_device(N, dev)
csr = CsrF64(ip, ix, dv, N, 0)
index = ShardedIndex(docs, 0, N, tfidf_local=csr)
Q = bench.make_queries(B, 1)[0].to(dev)
qc = bench.make_query_csr(B, dev)
def t(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): r = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, r
ms_d, (s, i) = t(lambda: index.search(Q, 50))
ms_t, tf = t(lambda: tfidf_candidates(i, csr, qc))
ms_r, out = t(lambda: hybrid_rerank(i, s, 0.5, tfidf=tf, top_n=10))
ms_h, _ = t(lambda: index.search_hybrid(Q, qc, 0.5, k=50, top_n=10))
print(f"N={N} B={B}: dense {ms_d:.3f} ms, tfidf_candidates {ms_t:.3f} ms, hybrid_rerank {ms_r:.3f} ms, search_hybrid {ms_h:.3f} ms")
