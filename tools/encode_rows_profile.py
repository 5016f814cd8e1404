"""Bulk encode through `encode_rows` (host rows -> resident matrix): passages/s with one and two compute lanes, SM clocks
under the sustained load, and the share of wall time the host spends packing.  usage: encode_rows_profile.py [n_passages]"""
import sys, time, numpy as np, torch
sys.path.insert(0, ".")
import bench
from twotowermlretrieval_b200 import TwoTowerModel, synth
from twotowermlretrieval_b200.encode import encode_rows

dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000
cfg = synth.default_config()
torch.manual_seed(0)
model = TwoTowerModel(cfg, None)
model.doc_encoder.embedding.weight.requires_grad_(False)
model.to(dev).eval()
enc = model.doc_encoder
flat, lens = synth.make_ragged_tokens(n, "passage", cfg["VOCAB_SIZE"], seed=2)
out = torch.empty(n, 256, device=dev)
encode_rows(enc, (flat[: int(lens[:4096].sum())], lens[:4096]), dev, out=out)
torch.cuda.synchronize()
for lanes, max_tokens, max_rows in ((1, 1048576, 30720), (1, 2097152, 61440), (1, 4194304, 122880), (1, 1048576, 30720)):
    smp = bench.ClockSampler(0); smp.start()
    t0 = time.perf_counter()
    encode_rows(enc, (flat, lens), dev, out=out, streams=lanes, max_tokens=max_tokens, max_rows=max_rows)
    t_host = time.perf_counter() - t0
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    clk = smp.stop()
    print(f"lanes {lanes} max_tokens {max_tokens} max_rows {max_rows}: {n} passages {int(lens.sum())} tokens: {dt*1e3:.1f} ms "
          f"({t_host*1e3:.1f} ms until the host returned) -> {n/dt:,.0f} passages/s, {lens.sum()/dt/1e6:.1f} Mtok/s; clocks {clk}")
