"""Small shapes of every tcgen05 / TMA / cluster kernel, for compute-sanitizer (racecheck, synccheck, memcheck).
Usage: compute-sanitizer --tool racecheck python tools/sanitize_driver.py [search|towers|train|all]"""
import sys, numpy as np, torch
sys.path.insert(0, ".")
from twotowermlretrieval_b200 import TwoTowerModel, synth, triplet_loss_cosine
from twotowermlretrieval_b200.index import search_topk
from twotowermlretrieval_b200.optim import FusedClipAdam
what = sys.argv[1] if len(sys.argv) > 1 else "all"
dev = torch.device("cuda:0")
if what in ("search", "all"):
    D = torch.tensor(synth.make_unit_rows(40_000, 256, seed=1), device=dev)
    for B in (2, 40, 200):                 # streaming kernel, tcgen05 single CTA (fused launch), CTA pairs
        Q = torch.tensor(synth.make_unit_rows(B, 256, seed=2 + B), device=dev)
        s, i = search_topk(Q, D, 50)
        torch.cuda.synchronize()
        print("search", B, float(s.sum()))
if what in ("towers", "train", "all"):
    cfg = synth.default_config(vocab_size=2000, embed_dim=200)
    cfg["DROPOUT"] = 0.0
    sd = synth.make_state_dict(cfg, seed=0, table_seed=1)
    m = TwoTowerModel(cfg, sd["query_encoder.embedding.weight"])
    m.load_state_dict({k: torch.tensor(v) for k, v in sd.items()})
    m.to(dev)
    ids, _ = synth.make_tokens(160, "passage", 2000, seed=12, lengths=np.random.default_rng(12).integers(4, 24, 160))
    x = torch.tensor(ids, device=dev)
    if what in ("towers", "all"):
        m.eval()
        with torch.no_grad():
            e = m.encode_document(x)      # fp16 inference pipeline: gather, kind::f16 projection, tcgen05 recurrence
        torch.cuda.synchronize()
        print("encode", float(e.sum()))
    if what in ("train", "all"):
        m.train()
        opt = FusedClipAdam(m, lr=1e-3, max_norm=1.0)
        opt.zero_grad()
        loss = triplet_loss_cosine((m.encode_query(x[:64, :8]), m.encode_document(x[:64]), m.encode_document(x[64:128])), margin=0.5)
        loss.backward()                   # tf32 projection, tcgen05 recurrence + BPTT, tf32 weight-gradient GEMMs
        opt.step()
        torch.cuda.synchronize()
        print("train", float(loss))
