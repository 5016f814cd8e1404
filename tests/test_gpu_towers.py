"""GPU parity: tower forward (gather -> tcgen05 projection -> GRU recurrence -> head) vs the
fixtures generated from the unmodified reference and vs the oracle (model.py:48-75)."""
import os

import numpy as np
import pytest
import torch

from conftest import golden_weights, load_golden
from gpu_util import model_from_numpy
from oracle import torch_path
from twotowermlretrieval_b200 import _lib, synth

pytestmark = pytest.mark.gpu

# north_star: embeddings within 1e-3 relative of the fp32 reference.  Rows are unit vectors
# (or O(0.1..1) un-normalised states), so the bound is applied as ||e - e_ref||_2 <= 1e-3 * ||e_ref||_2
# per row; the tf32 projection measures ~3e-4, the fp32 debug path ~1e-6.
REL_TOL = 1e-3


def assert_rows_close(got: torch.Tensor, want: np.ndarray, rel=REL_TOL):
    got = got.detach().cpu().double().numpy()
    want = np.asarray(want, dtype=np.float64)
    err = np.linalg.norm(got - want, axis=1) / np.maximum(np.linalg.norm(want, axis=1), 1e-12)
    assert err.max() <= rel, f"max relative row error {err.max():.3e} > {rel}"
    return float(err.max())


SMALL = ["small_bi2", "small_uni1", "small_bi1_trainable_table", "small_uni2"]


@pytest.mark.parametrize("name", SMALL)
def test_small_goldens_eval(cuda_device, name):
    g = load_golden(name)
    m = model_from_numpy(g["cfg"], golden_weights(g), cuda_device, pretrained=bool(g["pretrained"])).eval()
    with torch.no_grad():
        assert_rows_close(m.encode_query(torch.tensor(g["q"], device=cuda_device)), g["q_emb"])
        assert_rows_close(m.encode_document(torch.tensor(g["p"], device=cuda_device)), g["p_emb"])
        assert_rows_close(m.encode_document(torch.tensor(g["n"], device=cuda_device)), g["n_emb"])
        qe, de = m(torch.tensor(g["q"], device=cuda_device), torch.tensor(g["p"], device=cuda_device))
        assert_rows_close(qe, g["q_emb"]) and assert_rows_close(de, g["p_emb"])


def test_cfgdims_golden_cluster_kernel_and_tcgen05(cuda_device):
    g = load_golden("cfgdims")
    cfg = g["cfg"]
    m = model_from_numpy(cfg, synth.make_state_dict(cfg, seed=0, table_seed=1), cuda_device).eval()
    with torch.no_grad():
        e1 = assert_rows_close(m.encode_query(torch.tensor(g["q"], device=cuda_device)), g["q_emb"])
        e2 = assert_rows_close(m.encode_document(torch.tensor(g["p"], device=cuda_device)), g["p_emb"])
        e3 = assert_rows_close(m.encode_document(torch.tensor(g["n"], device=cuda_device)), g["n_emb"])
    print(f"\n[cfgdims] relative row error tf32 path: {max(e1, e2, e3):.3e}")


def test_cfgdims_fp32_debug_gemm_and_generic_gru_agree(cuda_device, monkeypatch):
    g = load_golden("cfgdims")
    cfg = g["cfg"]
    m = model_from_numpy(cfg, synth.make_state_dict(cfg, seed=0, table_seed=1), cuda_device).eval()
    x = torch.tensor(g["p"], device=cuda_device)
    monkeypatch.setenv("TTR_DEBUG_FP32_GEMM", "1")
    with torch.no_grad():
        # tcgen05 recurrence (fp16 matmul inputs, fp32 state) on top of the fp32 GEMM: its own error
        e = assert_rows_close(m.encode_document(x), g["p_emb"], rel=2e-4)
        print(f"\n[cfgdims] tcgen05 GRU + fp32 GEMM: relative row error {e:.3e}")
        for flag in (1024, 1):                                     # fp32 cluster GRU, generic GRU
            _lib.call_nostream("ttr_debug_set_flags", flag)
            try:
                assert_rows_close(m.encode_document(x), g["p_emb"], rel=2e-5)
            finally:
                _lib.call_nostream("ttr_debug_set_flags", 0)


def test_larger_batch_vs_oracle_at_config_dims(cuda_device):
    cfg = synth.default_config(vocab_size=30000, embed_dim=200)
    sd_np = synth.make_state_dict(cfg, seed=3, table_seed=4)
    m = model_from_numpy(cfg, sd_np, cuda_device).eval()
    ids, lens = synth.make_tokens(300, "passage", 30000, seed=21)
    qids, _ = synth.make_tokens(130, "query", 30000, seed=22)
    sd = torch_path.to_torch_state(sd_np)
    with torch.no_grad():
        ref_d = torch_path.encoder_forward(sd, "doc_encoder", torch.tensor(ids), cfg).numpy()
        ref_q = torch_path.encoder_forward(sd, "query_encoder", torch.tensor(qids), cfg).numpy()
        got_d = m.encode_document(torch.tensor(ids, device=cuda_device))
        got_q = m.encode_query(torch.tensor(qids, device=cuda_device))
    e = max(assert_rows_close(got_d, ref_d), assert_rows_close(got_q, ref_q))
    nrm = torch.linalg.vector_norm(got_d, dim=1)
    assert torch.allclose(nrm, torch.ones_like(nrm), atol=1e-5)          # query_inferencer.py:96 invariant
    print(f"\n[300 passages / 130 queries] relative row error: {e:.3e}")


def test_quirks_padding_order_and_zero_length(cuda_device):
    g = load_golden("small_bi2")
    m = model_from_numpy(g["cfg"], golden_weights(g), cuda_device).eval()
    dev = cuda_device
    with torch.no_grad():
        a = m.encode_query(torch.tensor([[5, 0, 7, 9]], device=dev))
        c = m.encode_query(torch.tensor([[5, 0, 7, 9, 0, 0, 0]], device=dev))
        b = m.encode_query(torch.tensor([[5, 0, 7, 0]], device=dev))
        assert torch.allclose(a, c, atol=1e-6)                            # padding invariance
        assert (a - b).abs().max() > 1e-4                                 # quirk #1: [5,0,7,0] has length 2
        x = torch.tensor(g["p"], device=dev)
        full = m.encode_document(x)
        perm = torch.randperm(x.shape[0], device=dev)
        assert torch.allclose(m.encode_document(x[perm]), full[perm], atol=1e-6)   # batch-order invariance
        one = torch.cat([m.encode_document(x[i:i + 1]) for i in range(x.shape[0])])
        assert torch.allclose(one, full, atol=1e-6)                       # batch-composition invariance
        bad = x.clone()
        bad[3] = 0
        with pytest.raises(RuntimeError, match="greater than 0"):
            m.encode_document(bad)                                        # quirk #2
        with pytest.raises(RuntimeError):
            m.encode_document(torch.zeros(2, 4, dtype=torch.long, device=dev))


def test_unnormalised_and_unidirectional(cuda_device):
    g = load_golden("small_uni1")
    assert not g["cfg"]["NORMALIZE_OUTPUT"] and not g["cfg"]["BIDIRECTIONAL"]
    m = model_from_numpy(g["cfg"], golden_weights(g), cuda_device).eval()
    with torch.no_grad():
        out = m.encode_document(torch.tensor(g["n"], device=cuda_device))
    assert_rows_close(out, g["n_emb"])
    assert (torch.linalg.vector_norm(out, dim=1) - 1).abs().max() > 1e-3


def test_tcgen05_gru_is_repeatable_and_matches_fp32_kernel(cuda_device, monkeypatch):
    """Multi-tile batch (two 256-row cluster tiles, both chains, partial last chain): the tcgen05
    recurrence must be bit-repeatable run to run (a protocol race shows up as run-to-run noise), in both
    storage pipelines, and — on the same fp32 gi — agree with the fp32 CUDA-core cluster kernel within the
    fp16-operand rounding."""
    cfg = synth.default_config(vocab_size=5000, embed_dim=200)
    m = model_from_numpy(cfg, synth.make_state_dict(cfg, seed=5, table_seed=6), cuda_device).eval()
    ids, _ = synth.make_tokens(500, "passage", 5000, seed=31)
    x = torch.tensor(ids, device=cuda_device)
    with torch.no_grad():
        runs16 = [m.encode_document(x) for _ in range(4)]                  # fp16-storage inference pipeline
        monkeypatch.setenv("TTR_FP32_PIPELINE", "1")
        runs = [m.encode_document(x) for _ in range(4)]                    # fp32 gi, tcgen05 recurrence
        _lib.call_nostream("ttr_debug_set_flags", 1024)
        try:
            ref = m.encode_document(x)                                     # fp32 gi, fp32 CUDA-core recurrence
        finally:
            _lib.call_nostream("ttr_debug_set_flags", 0)
    for r in runs[1:]:
        assert torch.equal(r, runs[0])
    for r in runs16[1:]:
        assert torch.equal(r, runs16[0])
    assert_rows_close(runs[0], ref.cpu().numpy(), rel=2e-4)
    assert_rows_close(runs16[0], ref.cpu().numpy(), rel=5e-4)


def test_fp16_storage_pipeline_vs_fp32_storage_pipeline(cuda_device, monkeypatch):
    """Inference stores X, gi and the inter-layer y as fp16 (fp32 accumulation/bias/state); the fp32-storage
    pipeline (tf32 projection, fp32 gi) is kept behind TTR_FP32_PIPELINE=1.  Both must meet the 1e-3 bound
    against the oracle, and differ from each other by the fp16 rounding of gi only."""
    cfg = synth.default_config(vocab_size=30000, embed_dim=200)
    cfg["DROPOUT"] = 0.0                                  # the training-mode comparison below must be deterministic
    sd_np = synth.make_state_dict(cfg, seed=3, table_seed=4)
    m = model_from_numpy(cfg, sd_np, cuda_device).eval()
    ids, _ = synth.make_tokens(300, "passage", 30000, seed=21)
    sd = torch_path.to_torch_state(sd_np)
    x = torch.tensor(ids, device=cuda_device)
    with torch.no_grad():
        ref = torch_path.encoder_forward(sd, "doc_encoder", torch.tensor(ids), cfg).numpy()
        e16 = m.encode_document(x)
        monkeypatch.setenv("TTR_FP32_PIPELINE", "1")
        e32 = m.encode_document(x)
    a, b = assert_rows_close(e16, ref), assert_rows_close(e32, ref)
    d = assert_rows_close(e16, e32.cpu().numpy(), rel=5e-4)
    print(f"\n[fp16 storage] vs oracle {a:.3e}; [fp32 storage] vs oracle {b:.3e}; between them {d:.3e}")
    # training mode keeps the fp32-storage path (autograd needs fp32 y / saved gates) and matches it bit for bit
    m.train()
    monkeypatch.delenv("TTR_FP32_PIPELINE")
    g = m.encode_document(x)
    assert g.requires_grad
    assert torch.equal(g.detach(), e32)


def test_bulk_encode_matches_per_batch_encode(cuda_device):
    from twotowermlretrieval_b200.encode import encode_rows
    cfg = synth.default_config(vocab_size=5000, embed_dim=200)
    sd_np = synth.make_state_dict(cfg, seed=5, table_seed=6)
    m = model_from_numpy(cfg, sd_np, cuda_device).eval()
    ids, lens = synth.make_tokens(500, "passage", 5000, seed=31)
    rows = [ids[i, :lens[i]].tolist() for i in range(500)]
    out = encode_rows(m.doc_encoder, rows, cuda_device, max_tokens=4096, max_rows=64)
    with torch.no_grad():
        ref = m.encode_document(torch.tensor(ids, device=cuda_device))
    assert torch.allclose(out, ref, atol=2e-6)
    # two compute lanes (default) give bit-identical rows to one, into a caller-owned matrix at an offset, and the
    # device-batch entry point agrees with plain forwards
    from twotowermlretrieval_b200.encode import encode_padded_batches
    one = encode_rows(m.doc_encoder, rows, cuda_device, max_tokens=4096, max_rows=64, streams=1)
    big = torch.zeros(520, cfg["HIDDEN_DIM"], device=cuda_device)
    encode_rows(m.doc_encoder, rows, cuda_device, out=big, out_offset=10, max_tokens=2048, max_rows=48, streams=3)
    assert torch.equal(out, one)
    assert torch.allclose(big[10:510], out, atol=2e-6) and float(big[:10].abs().sum()) == 0.0 and float(big[510:].abs().sum()) == 0.0
    order = np.argsort(-lens, kind="stable")
    bts = [torch.tensor(ids[order[i:i + 100], :int(lens[order[i]])], device=cuda_device) for i in range(0, 500, 100)]
    es = encode_padded_batches(m.doc_encoder, bts, streams=2)
    with torch.no_grad():
        for b, e in zip(bts, es):
            assert torch.equal(e, m.encode_document(b))
    with pytest.raises(RuntimeError):
        encode_rows(m.doc_encoder, rows[:3] + [[]], cuda_device)
    with pytest.raises(RuntimeError):                      # ids that are all 0 ("the"): zero effective length (quirk #2),
        encode_rows(m.doc_encoder, rows[:40] + [[0, 0, 0]] + rows[40:80], cuda_device)   # found by the packer's count
    # a zero in the middle of a row only shortens it (quirk #1); the packed token bound comes from the packer's count
    odd = [list(r) for r in rows[:64]]
    odd[5][1] = 0
    got = encode_rows(m.doc_encoder, odd, cuda_device, max_tokens=2048, max_rows=32)
    with torch.no_grad():
        T = max(len(r) for r in odd)
        pad = torch.tensor([r + [0] * (T - len(r)) for r in odd], device=cuda_device)
        assert torch.allclose(got, m.encode_document(pad), atol=2e-6)
