"""Projection GEMM timing probe: ttr_gemm_f16_bias at the encode shapes under the A/B debug flags, plus the device's
pure-write and copy bandwidth (the K = 200 layer writes 3,072 B and reads 400 B per token)."""
import sys, torch
sys.path.insert(0, ".")
from twotowermlretrieval_b200 import _lib
dev = torch.device("cuda:0")
M = int(sys.argv[1]) if len(sys.argv) > 1 else 504769

def timed(f, n=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

big = torch.empty(1 << 30, dtype=torch.float32, device=dev)          # 4 GiB
src = torch.empty(1 << 29, dtype=torch.float32, device=dev)
ms = timed(lambda: big.fill_(1.0)); print(f"fill_ 4 GiB: {ms:.3f} ms = {big.numel()*4/ms/1e6:.0f} GB/s write")
ms = timed(lambda: big[: 1 << 29].copy_(src)); print(f"copy_ 2 GiB: {ms:.3f} ms = {2*src.numel()*4/ms/1e6:.0f} GB/s read+write")
ms = timed(lambda: src.sum()); print(f"sum 2 GiB: {ms:.3f} ms = {src.numel()*4/ms/1e6:.0f} GB/s read")
del big, src
for K in (200, 512):
    X = (torch.randn(M, K, device=dev) * 0.3).half()
    W = (torch.randn(1536, K, device=dev) * 0.06).half()
    b = torch.zeros(1536, device=dev)
    C = torch.empty(M, 1536, dtype=torch.float16, device=dev)
    mv = torch.tensor([M], dtype=torch.int32, device=dev)
    for name, flags in (("default", 0), ("no stores (bit 11)", 2048), ("pair, W streamed (bit 19)", 1 << 19), ("single CTA (bit 30)", 1 << 30),
                        ("single CTA, no stores", (1 << 30) | 2048), ("coalesced-store epilogue (bit 18)", 1 << 18), ("clusters of 4, A multicast (bit 16)", 1 << 16)):
        _lib.call_nostream("ttr_debug_set_flags", flags)
        ms = timed(lambda: _lib.call("ttr_gemm_f16_bias", X, W, b, C, M, mv, 1536, K))
        _lib.call_nostream("ttr_debug_set_flags", 0)
        print(f"K={K} M={M} {name:28s}: {ms:.3f} ms = {2.0*M*K*1536/ms/1e9:7.1f} TFLOP/s; out {M*3072/ms/1e6:.0f} GB/s, in {M*K*2/ms/1e6:.0f} GB/s")
