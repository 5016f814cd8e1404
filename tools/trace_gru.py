"""Dump the per-step timeline of the tcgen05 GRU recurrence (cluster 0, CTA 0) — diagnostic."""
import sys, torch, numpy as np
sys.path.insert(0, ".")
from twotowermlretrieval_b200 import _lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
T = int(sys.argv[2]) if len(sys.argv) > 2 else 100
dev = torch.device("cuda:0")
H, dirs = 256, 2
torch.manual_seed(0)
lens = torch.full((B,), T, dtype=torch.int32)
offsets = torch.zeros(B + 1, dtype=torch.int32); offsets[1:] = torch.cumsum(lens, 0)
order = torch.arange(B, dtype=torch.int32)
tok = int(offsets[-1])
gi = torch.randn(tok, dirs * 3 * H, device=dev)
W = (torch.rand(dirs, 3 * H, H, device=dev) - 0.5) / 8
b = (torch.rand(dirs, 3 * H, device=dev) - 0.5) / 8
y = torch.empty(tok, dirs * H, device=dev); hl = torch.empty(B, dirs * H, device=dev)
offsets, order = offsets.to(dev), order.to(dev)
ws = torch.empty(int(_lib.load().ttr_gru_fwd_workspace_bytes(B, H, dirs)), dtype=torch.uint8, device=dev)
def run():
    _lib.call("ttr_gru_recurrence_fwd_ws", gi, W, b, order, offsets, B, H, dirs, y, hl, None, ws, ws.numel())
run(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
clusters = -(-B // 256) * dirs
print(f"B={B} T={T}: {ms:.3f} ms, {clusters} clusters; if 16 run at once: {ms * 1e3 / (T * -(-clusters // 16)):.2f} us per step")
tr = torch.zeros(8 * 256, dtype=torch.int64, device=dev)
_lib.call_nostream("ttr_debug_set_trace", tr.data_ptr())
run(); torch.cuda.synchronize()
_lib.call_nostream("ttr_debug_set_trace", None)
t = tr.cpu().numpy().reshape(8, 256)
names = ["step_start", "h_full", "mma_issued", "mma_done", "epi_done", "synced", "consumed", "copies_out"]
print("step " + " ".join(f"{n:>11}" for n in names))
for i in list(range(0, 4)) + list(range(40, 46)):
    print(f"{i:4d} " + " ".join(f"{int(t[r, i] - t[0, 0]):11d}" for r in range(8)))
sl = slice(10, min(T, 256) - 2)
print("cycles/step:", np.diff(t[0, sl]).mean())
for r in range(1, 8):
    print(f"  {names[r]:>11} - {names[r-1]:<11}: {(t[r, sl] - t[r-1, sl]).mean():8.0f}")
print(f"  next step_start - copies_out: {(t[0, 11:min(T,256)-1] - t[7, sl]).mean():8.0f}")
