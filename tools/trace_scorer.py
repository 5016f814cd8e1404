"""Dump the per-tile pipeline timeline of the tcgen05 scorer (CTA 0) — diagnostic."""
import sys, torch, numpy as np
sys.path.insert(0, ".")
from twotowermlretrieval_b200 import _lib
from twotowermlretrieval_b200.index import search_topk
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = torch.device("cuda:0")
D = torch.nn.functional.normalize(torch.randn(2_000_000, 256, device=dev), dim=1)
Q = torch.nn.functional.normalize(torch.randn(B, 256, device=dev), dim=1)
search_topk(Q, D, 50)
tr = torch.zeros(5 * 256, dtype=torch.int64, device=dev)
_lib.call_nostream("ttr_debug_set_trace", tr.data_ptr())
search_topk(Q, D, 50)
torch.cuda.synchronize()
_lib.call_nostream("ttr_debug_set_trace", None)
t = tr.cpu().numpy().reshape(5, 256)
t0 = t[0, 0]
names = ["prod_issue", "mma_full", "mma_issued", "epi_accfull", "epi_release"]
print("tile " + " ".join(f"{n:>12}" for n in names))
for i in list(range(0, 12)) + list(range(100, 112)):
    print(f"{i:4d} " + " ".join(f"{int(t[r, i] - t0):12d}" for r in range(5)))
d = np.diff(t[:, 20:250], axis=1)
print("mean cycles/tile per role:", d.mean(axis=1))
print("mma_full - prod_issue (TMA latency):", (t[1, 20:250] - t[0, 20:250]).mean())
print("mma_issued - mma_full:", (t[2, 20:250] - t[1, 20:250]).mean())
print("epi_accfull - mma_issued:", (t[3, 20:250] - t[2, 20:250]).mean())
print("epi_release - epi_accfull:", (t[4, 20:250] - t[3, 20:250]).mean())
