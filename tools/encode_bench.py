"""Doc-tower bulk encode micro-benchmark (config.json dims) with a per-kernel CUDA-event breakdown."""
import sys, time, numpy as np, torch
sys.path.insert(0, ".")
from twotowermlretrieval_b200 import TwoTowerModel, synth, _lib
from twotowermlretrieval_b200 import towers

import os
dev = torch.device("cuda:0")
if os.environ.get("TTR_DEBUG_FLAGS"):
    _lib.call_nostream("ttr_debug_set_flags", int(os.environ["TTR_DEBUG_FLAGS"]))
cfg = synth.default_config()
torch.manual_seed(0)
model = TwoTowerModel(cfg, None)
model.doc_encoder.embedding.weight.requires_grad_(False)
model.to(dev).eval()
model.doc_encoder.strict_lengths = False
NP = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
BS = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
ids, lens = synth.make_tokens(NP, "passage", cfg["VOCAB_SIZE"], seed=2)
order = np.argsort(-lens, kind="stable")
batches = [torch.tensor(ids[order[i:i + BS], :int(lens[order[i]])], device=dev) for i in range(0, NP, BS)]
toks = int(lens.sum())

# per-call timing via monkeypatched _lib.call
events = []
orig = _lib.call
def timed_call(name, *a):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); orig(name, *a); e1.record()
    events.append((name, e0, e1))
with torch.no_grad():
    for b in batches: model.encode_document(b)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        for b in batches: model.encode_document(b)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
    print(f"{NP} passages, {toks} tokens, batches of {BS}: {dt*1e3:.1f} ms -> {NP/dt:,.0f} passages/s, {toks/dt/1e6:.1f} Mtok/s")
    towers._lib.call = timed_call
    _lib.call = timed_call
    for b in batches: model.encode_document(b)
    torch.cuda.synchronize()
    _lib.call = orig; towers._lib.call = orig
agg = {}
for name, e0, e1 in events:
    agg.setdefault(name, []).append(e0.elapsed_time(e1))
tot = sum(sum(v) for v in agg.values())
for name, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"  {name:28s} calls {len(v):3d} total {sum(v):8.2f} ms  ({100*sum(v)/tot:5.1f} %)")
# projection GEMM utilisation: per layer FLOPs
g = [e0.elapsed_time(e1) for n, e0, e1 in events if n in ("ttr_gemm_tf32_bias", "ttr_gemm_f16_bias")]
kind = "fp16" if any(n == "ttr_gemm_f16_bias" for n, _, _ in events) else "tf32"
bpe = 2 if kind == "fp16" else 4
nb = len(batches)
l0 = sum(g[0::2]); l1 = sum(g[1::2])
f0 = 2 * toks * 200 * 1536; f1 = 2 * toks * 512 * 1536
print(f"  input projection L0: {l0:.2f} ms = {f0/l0/1e9:.1f} TFLOP/s ; L1: {l1:.2f} ms = {f1/l1/1e9:.1f} TFLOP/s ({kind}; valid tokens only)")
print(f"  output bytes gi per layer: {toks*1536*bpe/1e9:.2f} GB -> L0 {toks*1536*bpe/l0/1e6:.0f} GB/s, L1 {toks*1536*bpe/l1/1e6:.0f} GB/s write")
