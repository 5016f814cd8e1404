"""Training-step micro-benchmark (config.json dims, BASELINE config 5 per-GPU shape: 1024 triplets)."""
import sys, time, numpy as np, torch
sys.path.insert(0, ".")
from twotowermlretrieval_b200 import TwoTowerModel, synth, _lib, triplet_loss_cosine
from twotowermlretrieval_b200 import towers, towers_bwd, optim
from twotowermlretrieval_b200.optim import FusedClipAdam

dev = torch.device("cuda:0")
cfg = synth.default_config()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
torch.manual_seed(0)
model = TwoTowerModel(cfg, None)
for enc in (model.query_encoder, model.doc_encoder):
    enc.embedding.weight.requires_grad_(False)
    enc.strict_lengths = False
model.to(dev).train()
opt = FusedClipAdam(model, lr=cfg["LR"], max_norm=1.0)
q, ql = synth.make_tokens(B, "query", cfg["VOCAB_SIZE"], seed=2)
p, pl = synth.make_tokens(B, "passage", cfg["VOCAB_SIZE"], seed=3)
n, nl = synth.make_tokens(B, "passage", cfg["VOCAB_SIZE"], seed=4)
qd, pd_, nd = (torch.tensor(a, device=dev) for a in (q, p, n))
toks = int(ql.sum() + pl.sum() + nl.sum())

from twotowermlretrieval_b200.trainer import TwoTowerTrainer
LANES = int(sys.argv[2]) if len(sys.argv) > 2 else 3
trainer = TwoTowerTrainer(model, opt, lambda tr: triplet_loss_cosine(tr, margin=cfg["MARGIN"]), dev, dict(cfg, TOWER_STREAMS=LANES))
print(f"tower streams: {LANES}")

def step():
    return trainer.train_step(qd, pd_, nd)[0]

for _ in range(2): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3): step()
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 3
print(f"B={B} triplets, {toks} tokens/step (padded ids [{q.shape[1]},{p.shape[1]},{n.shape[1]}]): {dt*1e3:.1f} ms/step -> {B/dt:,.0f} triplets/s; mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB")
events = []
orig = _lib.call
def timed_call(name, *a):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); orig(name, *a); e1.record()
    events.append((name, e0, e1))
for mod in (towers, towers_bwd, optim): mod._lib.call = timed_call
_lib.call = timed_call
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); step(); e1.record()
torch.cuda.synchronize()
_lib.call = orig
for mod in (towers, towers_bwd, optim): mod._lib.call = orig
agg = {}
for name, a, b in events: agg.setdefault(name, []).append(a.elapsed_time(b))
tot = e0.elapsed_time(e1)
ksum = sum(sum(v) for v in agg.values())
print(f"  instrumented step {tot:.1f} ms, inside C-ABI calls {ksum:.1f} ms")
for name, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"  {name:28s} calls {len(v):3d} total {sum(v):8.2f} ms  ({100*sum(v)/tot:5.1f} %)")
