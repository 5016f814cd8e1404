"""Bring-up check of the CTA-pair scorer (B > 128): default vs one-CTA-per-query-tile (debug bit 27) vs swapped
half assignment (bit 26), with index-difference statistics — tells a wrong half convention (ids off by 32) from
anything else."""
import sys, torch
sys.path.insert(0, ".")
from twotowermlretrieval_b200 import _lib, synth
from twotowermlretrieval_b200.index import search_topk
dev = torch.device("cuda:0")
for B, N in ((256, 100_000), (200, 5_000), (512, 300_000), (256, 1_000_000)):
    D = torch.tensor(synth.make_unit_rows(N, 256, seed=1), device=dev)
    Q = torch.tensor(synth.make_unit_rows(B, 256, seed=2), device=dev)
    res = {}
    for name, fl in (("pair", 0), ("single", 1 << 27), ("pair_swapped", 1 << 26), ("pair_twopass", 1 << 21)):
        _lib.call_nostream("ttr_debug_set_flags", fl)
        try:
            s, i = search_topk(Q, D, 50)
            torch.cuda.synchronize()
            res[name] = (s.clone(), i.clone())
        except Exception as e:
            res[name] = None
            print(B, N, name, "FAILED:", repr(e)[:300])
        finally:
            _lib.call_nostream("ttr_debug_set_flags", 0)
    ref = res["single"]
    for name in ("pair", "pair_swapped", "pair_twopass"):
        if res[name] is None or ref is None:
            continue
        s, i = res[name]
        d = (i - ref[1])
        print(f"B={B} N={N} {name}: idx equal {float((d == 0).float().mean()):.4f}, |diff|==32 {float((d.abs() == 32).float().mean()):.4f}, "
              f"scores equal {float((s == ref[0]).float().mean()):.4f}, max score diff {float((s - ref[0]).abs().max()):.3e}")
