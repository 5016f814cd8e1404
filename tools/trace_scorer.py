"""Dump the per-tile pipeline timeline of the tcgen05 scorer (CTA 0) — diagnostic."""
import sys, torch, numpy as np
sys.path.insert(0, ".")
from twotowermlretrieval_b200 import _lib
from twotowermlretrieval_b200.index import search_topk
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = torch.device("cuda:0")
N = int(sys.argv[3]) if len(sys.argv) > 3 else 2_000_000
D = torch.empty(N, 256, device=dev)
for lo_ in range(0, N, 1 << 20):
    hi_ = min(N, lo_ + (1 << 20))
    D[lo_:hi_] = torch.nn.functional.normalize(torch.randn(hi_ - lo_, 256, device=dev), dim=1)
Q = torch.nn.functional.normalize(torch.randn(B, 256, device=dev), dim=1)
search_topk(Q, D, 50)
tr = torch.zeros(8 * 256, dtype=torch.int64, device=dev)
flags = int(sys.argv[2]) if len(sys.argv) > 2 else 0
_lib.call_nostream("ttr_debug_set_flags", flags)
_lib.call_nostream("ttr_debug_set_trace", tr.data_ptr())
search_topk(Q, D, 50)
torch.cuda.synchronize()
_lib.call_nostream("ttr_debug_set_trace", None)
_lib.call_nostream("ttr_debug_set_flags", 0)
t = tr.cpu().numpy().reshape(8, 256)
t0 = t[0, 0]
mk = t[:5, 255]
print("kernel marks (cycles from entry): prologue_done %d, first_tma_issue %d, last_tile_filtered %d, published %d, exit %d" %
      (mk[1] - mk[0], t0 - mk[0], mk[2] - mk[0], mk[3] - mk[0], mk[4] - mk[0]))
ntile = -(-N // 32 + 147) // 148
print("tiles per CTA ~", ntile, "; cycles from first TMA issue to last filtered:", mk[2] - t0)
names = ["prod_issue", "mma_full", "mma_issued", "epi_accfull", "epi_release", "mma_done", "epi_loop_end", "epi_tile_end"]
print("tile " + " ".join(f"{n:>12}" for n in names))
rows = range(0, 256) if (len(sys.argv) > 4 and sys.argv[4] == "all") else list(range(0, 12)) + list(range(100, 112))
for i in rows:
    print(f"{i:4d} " + " ".join(f"{int(t[r, i] - t0):12d}" for r in range(8)))
d = np.diff(t[:, 20:250], axis=1)
print("mean cycles/tile per role:", d.mean(axis=1))
print("mma_full - prod_issue (TMA latency):", (t[1, 20:250] - t[0, 20:250]).mean())
print("mma_issued - mma_full:", (t[2, 20:250] - t[1, 20:250]).mean())
print("epi_accfull - mma_issued:", (t[3, 20:250] - t[2, 20:250]).mean())
print("epi_release - epi_accfull:", (t[4, 20:250] - t[3, 20:250]).mean())
if flags & 64:
    print("mma_done - mma_issued (in-kernel MMA execution of one 32-MMA tile):", (t[5, 20:250] - t[2, 20:250]).mean())
    print("mma_done - mma_full:", (t[5, 20:250] - t[1, 20:250]).mean())
print("epi_loop_end - epi_release (compare/append loop):", (t[6, 20:250] - t[4, 20:250]).mean())
print("epi_tile_end - epi_loop_end (vote + compaction):", (t[7, 20:250] - t[6, 20:250]).mean(), "max", (t[7, 20:250] - t[6, 20:250]).max())
print("next epi_accfull - epi_tile_end (tau refresh + wait):", (t[3, 21:251] - t[7, 20:250]).mean())
