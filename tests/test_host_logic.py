"""CPU: host-side logic — state_dict/seed compatibility with the reference module, flat
parameter views, shard bounds, bulk-encode batching, tokenizer, 2-rank gloo merge scheme."""
import os
import pickle
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

from conftest import golden_weights, load_golden
from oracle import towers_numpy as onp
from twotowermlretrieval_b200 import TwoTowerModel, synth
from twotowermlretrieval_b200.encode import plan_batches
from twotowermlretrieval_b200.index import shard_bounds
from twotowermlretrieval_b200.tokenizer import PretrainedTokenizer

REF = Path("/root/reference/backend")


def test_state_dict_keys_and_shapes_match_reference_contract():
    cfg = synth.default_config(vocab_size=300, embed_dim=200)
    m = TwoTowerModel(cfg)
    sd = m.state_dict()
    want = {}
    for tower in ("query_encoder", "doc_encoder"):
        for k, shape in synth.tower_param_shapes(cfg).items():
            want[f"{tower}.{k}"] = shape
    assert set(sd.keys()) == set(want.keys())
    for k, shape in want.items():
        assert tuple(sd[k].shape) == tuple(shape), k
    n_train = sum(p.numel() for n, p in m.named_parameters() if "embedding" not in n)
    assert n_train == 4035072          # trainable GRU/projection params of both towers


@pytest.mark.skipif(not REF.exists(), reason="reference tree not present (GPU box)")
def test_same_seed_gives_reference_initialisation_and_loads_both_ways():
    sys.path.insert(0, str(REF))
    import model as refmodel
    cfg = synth.default_config(vocab_size=120, embed_dim=20)
    cfg["HIDDEN_DIM"] = 24
    torch.manual_seed(123)
    ref = refmodel.TwoTowerModel(cfg)
    torch.manual_seed(123)
    ours = TwoTowerModel(cfg)
    rs, os_ = ref.state_dict(), ours.state_dict()
    assert list(rs.keys()) == list(os_.keys())
    for k in rs:
        assert torch.equal(rs[k], os_[k]), k
    ours.load_state_dict(rs)
    ref.load_state_dict(ours.state_dict())
    table = np.random.default_rng(0).standard_normal((120, 20)).astype(np.float32)
    ours2 = TwoTowerModel(cfg, table)
    assert not ours2.query_encoder.embedding.weight.requires_grad     # model.py:25-27


def test_flat_views_survive_load_and_move():
    cfg = synth.default_config(vocab_size=60, embed_dim=12)
    cfg["HIDDEN_DIM"] = 16
    m = TwoTowerModel(cfg)
    flat = m.flat_params()
    sd = {k: torch.randn_like(v) for k, v in m.state_dict().items()}
    m.load_state_dict(sd)
    assert m._views_ok()
    w, b, whh, bhh = m.query_encoder.layer_weights(1)
    assert torch.equal(w[:48], sd["query_encoder.rnn.weight_ih_l1"])
    assert torch.equal(w[48:], sd["query_encoder.rnn.weight_ih_l1_reverse"])
    assert torch.equal(whh[1], sd["query_encoder.rnn.weight_hh_l1_reverse"])
    assert torch.equal(bhh[48:], sd["query_encoder.rnn.bias_hh_l1_reverse"])
    m.double().float()                      # _apply re-allocates storage -> lazily re-packed
    assert m.flat_params().numel() == flat.numel() and m._views_ok()
    g = m.flat_grads()
    assert all(p.grad is not None and p.grad.data_ptr() >= g.data_ptr() for n, p in m.named_parameters()
               if "embedding" not in n)


def test_lstm_is_rejected():
    with pytest.raises(NotImplementedError):
        TwoTowerModel(dict(synth.default_config(50, 8), RNN_TYPE="LSTM"))


def test_shard_bounds_cover_rows_exactly_once():
    for n in (0, 1, 7, 8, 1000, 8841823):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert all(hi - lo <= -(-n // world) for lo, hi in spans)


def test_plan_batches_is_a_partition_within_budget():
    rng = np.random.default_rng(0)
    lengths = synth.make_lengths(5000, "passage", rng)
    order, bounds = plan_batches(lengths, max_tokens=8192, max_rows=256)
    assert sorted(order.tolist()) == list(range(5000))
    assert bounds[0][0] == 0 and bounds[-1][1] == 5000
    for lo, hi in bounds:
        T = lengths[order[lo]]
        assert hi - lo <= 256 and ((hi - lo) * T <= 8192 or hi - lo == 1)
        assert (lengths[order[lo:hi]] <= T).all()


def test_tokenizer_matches_reference_rules(tmp_path):
    w2i = {"the": 0, "cat": 1, ".": 2, "sat": 3}
    p = tmp_path / "w.pkl"
    with open(p, "wb") as f:
        pickle.dump(w2i, f)
    tok = PretrainedTokenizer(str(p))
    assert tok.vocab_size() == 5 and tok.unk_token_id == 4                   # tokenizer.py:20-24
    assert tok.encode("The CAT sat, on... the mat!") == [0, 1, 3, 4, 4, 2, 2, 2, 0, 4, 4]
    assert tok.encode("") == [] and tok.decode([1, 3, 99]) == "cat sat <UNK>"
    if REF.exists():
        sys.path.insert(0, str(REF))
        import tokenizer as reftok
        rt = reftok.PretrainedTokenizer(str(p))
        for s in ["The CAT sat, on... the mat!", "a_b c-d; e?f", "Ünïcode wörds 123", ""]:
            assert rt.encode(s) == tok.encode(s)
    with pytest.raises(FileNotFoundError):
        PretrainedTokenizer(str(tmp_path / "missing.pkl"))


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        D = synth.make_unit_rows(4001, 256, seed=3)
        Q = synth.make_unit_rows(6, 256, seed=4)
        lo, hi = shard_bounds(D.shape[0], world, rank)
        s, i = onp.cosine_topk(Q, D[lo:hi], 50, dtype=np.float32)      # stand-in for the local kernel
        s, i = torch.tensor(s), torch.tensor(i + lo)
        gs = [torch.empty_like(s) for _ in range(world)]
        gi = [torch.empty_like(i) for _ in range(world)]
        dist.all_gather(gs, s)
        dist.all_gather(gi, i)
        cs, ci = torch.cat(gs, 1).numpy(), torch.cat(gi, 1).numpy()
        order = np.lexsort((ci, -cs), axis=1)[:, :50]
        ms, mi = np.take_along_axis(cs, order, 1), np.take_along_axis(ci, order, 1)
        fs, fi = onp.cosine_topk(Q, D, 50, dtype=np.float32)
        q.put((rank, bool((mi == fi).all()), float(np.abs(ms - fs).max())))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_shard_gather_merge_equals_global_topk():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = [q.get(timeout=180) for _ in range(2)]
    [p.join(60) for p in procs]
    assert all(ok for _, ok, _ in res) and all(err < 1e-6 for _, _, err in res)


def test_corpus_artifact_reader_formats_and_errors(tmp_path):
    """`artifacts.load_corpus_artifacts` reads the reference writer's files (`backend/main.py:134-149`) without a
    GPU: documents.pkl, document_embeddings.npy (fp32, memory-mapped), tfidf_artifacts.pkl; inconsistent
    directories and missing files are reported like the reference's readers would (`open` -> FileNotFoundError)."""
    import pickle
    from sklearn.feature_extraction.text import TfidfVectorizer
    from twotowermlretrieval_b200.artifacts import load_corpus_artifacts
    from twotowermlretrieval_b200.data import pad_rows
    docs = ["deep neural network layer", "vector search index", "machine learning model for text", "image and video"]
    vec = TfidfVectorizer(stop_words="english", max_features=20000)
    mat = vec.fit_transform(docs)
    emb = np.arange(len(docs) * 8, dtype=np.float32).reshape(len(docs), 8)
    with open(tmp_path / "documents.pkl", "wb") as fh:
        pickle.dump(docs, fh)
    np.save(tmp_path / "document_embeddings.npy", emb)
    with open(tmp_path / "tfidf_artifacts.pkl", "wb") as fh:
        pickle.dump({"vectorizer": vec, "matrix": mat}, fh)
    d2, e2, v2, m2 = load_corpus_artifacts(tmp_path)
    assert d2 == docs and np.array_equal(np.asarray(e2), emb) and e2.dtype == np.float32
    assert (m2 != mat).nnz == 0 and v2.vocabulary_ == vec.vocabulary_
    np.save(tmp_path / "document_embeddings.npy", emb[:3])
    with pytest.raises(ValueError):
        load_corpus_artifacts(tmp_path)
    with pytest.raises(FileNotFoundError):
        load_corpus_artifacts(tmp_path / "nope")
    p = pad_rows([[1, 2, 3], [], [7]])
    assert p.dtype == torch.int64 and p.tolist() == [[1, 2, 3], [0, 0, 0], [7, 0, 0]]


def _gloo_worker2(rank, world, port, q):
    """Host-side N > 1 logic on CPU: (a) the peer-exchange decision is AGREED across ranks — a rank whose symmetric
    allocation fails takes every rank to the all-gather path instead of splitting the group;
    (b) differing batch sizes raise on every rank; (c) bench.py's verifier (fp64 top-k over all shards,
    rank-identity check) agrees with the unsharded oracle."""
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    out = {}
    try:
        from twotowermlretrieval_b200 import _lib, index as ix

        class Boom:
            def __init__(self, group, device, max_b, k):
                if dist.get_rank() == 1:
                    raise RuntimeError("no P2P mapping on this rank")
                self.k, self.max_b = k, max_b

        real = ix._PeerExchange
        ix._PeerExchange = Boom
        try:
            idx = ix.ShardedIndex.__new__(ix.ShardedIndex)
            idx.docs = torch.zeros(4, 256)
            idx.group, idx._dist, idx.world, idx.rank = None, dist, world, rank
            idx.peer_memory, idx._px = True, None
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                out["agreed_fallback"] = idx._exchange(8, 50) is None and idx.peer_memory is False
            idx.peer_memory, idx._px = True, None
            try:
                idx._exchange(8 + rank, 50)          # B differs across ranks
                out["mismatch_raises"] = False
            except _lib.TTRError:
                out["mismatch_raises"] = True
        finally:
            ix._PeerExchange = real
        # (c) the bench verifier on CPU tensors
        import bench
        c = bench.Ctx()
        c.world, c.rank, c.dist, c.dev = world, rank, dist, torch.device("cpu")
        D = torch.tensor(synth.make_unit_rows(5003, 256, seed=3))
        Q = torch.tensor(synth.make_unit_rows(4, 256, seed=4))
        lo, hi = shard_bounds(D.shape[0], world, rank)
        s, i = bench.fp64_topk_all_shards(c, Q, D[lo:hi], lo, 50)
        fs, fi = onp.cosine_topk(Q.numpy(), D.numpy(), 50, dtype=np.float64)
        out["verifier_topk"] = bool((i.numpy() == fi).all()) and float(np.abs(s.numpy() - fs).max()) < 1e-12
        good = bench.check_against_fp64(c, Q, s.float(), i, D[lo:hi], lo, 50)
        bad_i = i.clone()
        bad_i[0, 3] = (bad_i[0, 3] + 7) % 5003       # a wrong document must be caught
        bad = bench.check_against_fp64(c, Q, s.float(), bad_i, D[lo:hi], lo, 50)
        out["verifier_accepts"], out["verifier_rejects"] = good["ok"], not bad["ok"]
        out["same_true"] = bench.same_on_all_ranks(c, i)
        out["same_false"] = not bench.same_on_all_ranks(c, i + rank)
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_exchange_agreement_and_bench_verifier():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker2, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = dict(q.get(timeout=300) for _ in range(2))
    [p.join(60) for p in procs]
    for rank in (0, 1):
        assert all(res[rank].values()), (rank, res[rank])


def test_host_packer_matches_pad_sequence():
    """`ttr_pack_padded_i64` (host C helper of encode_rows) == `pad_sequence(batch_first=True, padding_value=0)`
    (backend/main.py:50-56) on the selected rows."""
    from twotowermlretrieval_b200 import _lib
    rng = np.random.default_rng(3)
    lengths = rng.integers(1, 40, size=200).astype(np.int64)
    flat = rng.integers(1, 1000, size=int(lengths.sum())).astype(np.int64)
    starts = np.zeros(201, dtype=np.int64)
    np.cumsum(lengths, out=starts[1:])
    rows = np.ascontiguousarray(np.argsort(-lengths, kind="stable")[:64].astype(np.int64))
    T = int(lengths[rows[0]])
    out = torch.full((64, T), -1, dtype=torch.int64)
    _lib.call_nostream("ttr_pack_padded_i64", flat.ctypes.data, starts.ctypes.data, lengths.ctypes.data, rows.ctypes.data,
                       64, T, out.data_ptr())
    ref = torch.nn.utils.rnn.pad_sequence([torch.tensor(flat[starts[r]:starts[r] + lengths[r]]) for r in rows],
                                          batch_first=True, padding_value=0)
    assert torch.equal(out, ref)
    # the counting variant: same matrix + number of non-zero ids + number of rows without any
    import ctypes
    flat2 = flat.copy()
    flat2[starts[rows[3]]:starts[rows[3]] + lengths[rows[3]]] = 0          # an all-zero row
    flat2[starts[rows[5]] + 1] = 0                                         # a zero in the middle of a row
    out2 = torch.full((64, T), -1, dtype=torch.int64)
    nnz, zero = ctypes.c_int64(-1), ctypes.c_int64(-1)
    _lib.call_nostream("ttr_pack_padded_count_i64", flat2.ctypes.data, starts.ctypes.data, lengths.ctypes.data,
                       rows.ctypes.data, 64, T, out2.data_ptr(), ctypes.addressof(nnz), ctypes.addressof(zero))
    ref2 = torch.nn.utils.rnn.pad_sequence([torch.tensor(flat2[starts[r]:starts[r] + lengths[r]]) for r in rows],
                                           batch_first=True, padding_value=0)
    assert torch.equal(out2, ref2)
    assert nnz.value == int((ref2 != 0).sum()) and zero.value == 1


def test_plan_batches_rounds_rows_to_whole_recurrence_waves():
    from twotowermlretrieval_b200.encode import ROW_QUANTUM
    rng = np.random.default_rng(1)
    lengths = synth.make_lengths(60000, "passage", rng)
    order, bounds = plan_batches(lengths, max_tokens=524288, max_rows=15360)
    assert bounds[0][0] == 0 and bounds[-1][1] == 60000
    for lo, hi in bounds[:-1]:
        assert (hi - lo) % ROW_QUANTUM == 0 or hi - lo < ROW_QUANTUM


def test_plan_batches_radix_order_equals_the_stable_descending_sort():
    """`plan_batches` sorts 16-bit keys (numpy's radix path) when every length fits; the order must be the stable
    descending-length order of the general path, and lengths >= 65536 must fall back to it."""
    rng = np.random.default_rng(5)
    for hi in (40, 300, 70000):
        lengths = rng.integers(1, hi, size=5000).astype(np.int64)
        order, bounds = plan_batches(lengths, max_tokens=1 << 20, max_rows=30720)
        np.testing.assert_array_equal(order, np.argsort(-lengths, kind="stable"))
        assert bounds[0][0] == 0 and bounds[-1][1] == 5000 and all(a[1] == b[0] for a, b in zip(bounds, bounds[1:]))
        for lo, hi_ in bounds:                       # a batch never exceeds the padded-token budget (one row always fits)
            assert (hi_ - lo) * int(lengths[order[lo]]) <= (1 << 20) or hi_ - lo == 1


def test_bench_reads_roofline_traffic_from_the_committed_capture_and_says_so():
    import bench
    t = bench.committed_traffic(8841823, 128, 1)
    assert t["traffic"] is not None and 0.99 < t["traffic"] / (8841823 * 1024) < 1.02      # DRAM bytes == algorithmic bytes
    assert "COMMITTED CAPTURE" in t["traffic_note"] and "not measured in this run" in t["traffic_note"]
    assert bench.committed_traffic(8841823, 128, 8)["traffic"] is None                     # another shape: null, not a guess
    assert bench.committed_traffic(12345, 7, 1)["traffic"] is None
