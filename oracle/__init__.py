"""CPU oracle for the TwoTowerMLRetrieval hot path — TEST INFRASTRUCTURE ONLY.

Nothing under `twotowermlretrieval_b200/` imports this package.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference` legs
may import it, and there only as the checker / the reported CPU baseline.

What it restates (all citations are into the read-only reference tree):

* `oracle/towers_numpy.py`  — explicit-loop numpy (float64 or float32) restatement of
  `RNNEncoder.forward` (`backend/model.py:48-75`), `triplet_loss_cosine`
  (`backend/model.py:109-114`), dense scoring + top-k (`backend/evaluators.py:185-186`),
  the frontend hybrid rerank (`frontend/main.py:162-198`), the corpus-wide hybrid search
  (`backend/simple_hybrid.py:45-66`), `clip_grad_norm_` + Adam (`backend/main.py:257-259`).
* `oracle/torch_path.py`    — the same path through the same third-party call sites the
  reference uses (torch `nn.GRU` on packed sequences, `F.normalize`,
  `F.cosine_similarity`, `torch.matmul`/`torch.topk`, `torch.optim.Adam`), functional over
  a state_dict, with autograd; this is the "port" timed as `cpu_baseline`.

The arithmetic of the path lives in third-party PyTorch (`torch>=2.0.0`,
`requirements.txt:4`, un-pinned; 2.11.0+cu128 in this image) and scikit-learn
(`requirements.txt:6`; 1.9.0 here).  The reference ships no tests or golden vectors
(SURVEY.md §4), so parity is pinned the other way the task allows: `oracle/make_golden.py`
imports the UNMODIFIED `/root/reference/backend/model.py` in the build container, runs it
on seeded inputs and commits the outputs under `tests/golden/`; `tests/test_oracle.py`
checks both restatements against those fixtures on every CPU run.
"""
