import sys, numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from twotowermlretrieval_b200 import synth, _lib
from twotowermlretrieval_b200.encode import encode_rows
from gpu_util import model_from_numpy
dev = torch.device("cuda:0")
cfg = synth.default_config(vocab_size=5000, embed_dim=200)
sd_np = synth.make_state_dict(cfg, seed=5, table_seed=6)
m = model_from_numpy(cfg, sd_np, dev).eval()
ids, lens = synth.make_tokens(500, "passage", 5000, seed=31)
rows = [ids[i, :lens[i]].tolist() for i in range(500)]
def both():
    out = encode_rows(m.doc_encoder, rows, dev, max_tokens=4096, max_rows=64)
    with torch.no_grad():
        ref = m.encode_document(torch.tensor(ids, device=dev))
    return out, ref
for flag in (0, 1024):
    _lib.call_nostream("ttr_debug_set_flags", flag)
    o1, r1 = both(); o2, r2 = both()
    d = (o1 - r1).abs()
    rowerr = torch.linalg.vector_norm(o1 - r1, dim=1)
    print(f"flag {flag}: repeat-determinism out {float((o1-o2).abs().max()):.3e} ref {float((r1-r2).abs().max()):.3e}; bulk-vs-full max abs {float(d.max()):.3e}, max row err {float(rowerr.max()):.3e}, rows differing {int((rowerr>0).sum())}")
_lib.call_nostream("ttr_debug_set_flags", 0)
