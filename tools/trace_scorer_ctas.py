"""Per-CTA entry/exit times (globaltimer) of the main scoring kernel — straggler diagnostic."""
import sys, torch, numpy as np
sys.path.insert(0, ".")
from twotowermlretrieval_b200 import _lib
from twotowermlretrieval_b200.index import search_topk
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1105228
dev = torch.device("cuda:0")
D = torch.empty(N, 256, device=dev)
for lo_ in range(0, N, 1 << 20):
    hi_ = min(N, lo_ + (1 << 20))
    D[lo_:hi_] = torch.nn.functional.normalize(torch.randn(hi_ - lo_, 256, device=dev), dim=1)
Q = torch.nn.functional.normalize(torch.randn(B, 256, device=dev), dim=1)
for _ in range(3): search_topk(Q, D, 50)
tr = torch.zeros(8 * 256, dtype=torch.int64, device=dev)
_lib.call_nostream("ttr_debug_set_flags", 1 << 20)
_lib.call_nostream("ttr_debug_set_trace", tr.data_ptr())
search_topk(Q, D, 50); torch.cuda.synchronize()
_lib.call_nostream("ttr_debug_set_trace", None); _lib.call_nostream("ttr_debug_set_flags", 0)
t = tr.cpu().numpy().reshape(8, 256)
ent, ext = t[5, :148].astype(np.float64), t[6, :148].astype(np.float64)
t0 = ent.min()
print(f"N={N} B={B}: entry spread {ent.max()-t0:.0f} ns; exit min {ext.min()-t0:.0f} median {np.median(ext)-t0:.0f} max {ext.max()-t0:.0f} ns")
dur = ext - ent
print(f"CTA duration: min {dur.min():.0f} median {np.median(dur):.0f} p90 {np.percentile(dur,90):.0f} max {dur.max():.0f} ns")
order = np.argsort(-dur)[:8]
print("slowest CTAs (slice, duration ns):", [(int(i), int(dur[i])) for i in order])
