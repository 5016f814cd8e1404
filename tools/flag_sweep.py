"""Timing experiment: tcgen05 scorer with parts of the pipeline disabled (results invalid)."""
import sys, torch
sys.path.insert(0, ".")
from twotowermlretrieval_b200 import _lib
from twotowermlretrieval_b200.index import search_topk
dev = torch.device("cuda:0")
N = 8_841_823
D = torch.empty(N, 256, device=dev)
for lo in range(0, N, 1 << 20):
    hi = min(N, lo + (1 << 20))
    D[lo:hi] = torch.nn.functional.normalize(torch.randn(hi - lo, 256, device=dev), dim=1)
for B in (16, 128, 1024):
    Q = torch.nn.functional.normalize(torch.randn(B, 256, device=dev), dim=1)
    for flags, name in ((0, "normal"), (8, "no LDTM"), (16, "B from stage 0"), (48, "B pinned to stage0/kblock0"), (56, "no LDTM + B pinned")):
        _lib.call_nostream("ttr_debug_set_flags", flags)
        for _ in range(2):
            search_topk(Q, D, 50)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            search_topk(Q, D, 50)
        e1.record(); torch.cuda.synchronize()
        print(f"B={B:5d} {name:28s} {e0.elapsed_time(e1)/5:8.3f} ms")
    _lib.call_nostream("ttr_debug_set_flags", 0)
