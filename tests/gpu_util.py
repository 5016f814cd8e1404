"""Helpers shared by the -m gpu parity tests."""
import numpy as np
import torch

from oracle import towers_numpy as onp


def check_topk(scores, idx, Q, D, k, row_offset=0, rtol=1e-3, atol=1e-4):
    """Compare a GPU top-k with the fp64 oracle: scores within `rtol` relative (north_star:
    <= 1e-3 vs fp32) with an absolute floor of 1e-4 on the cosine scale [-1, 1] (the tcgen05
    path rounds operands to tf32, 2^-11 per element, so near-zero scores keep an absolute
    error of a few 1e-5); indices identical except where the oracle scores tie inside that
    tolerance at the position or at the k-th boundary."""
    s_ref, i_ref = onp.cosine_topk(Q, D, k, dtype=np.float64)
    scores, idx = scores.cpu().numpy(), idx.cpu().numpy() - row_offset
    kk = s_ref.shape[1]
    assert np.allclose(scores[:, :kk], s_ref, rtol=rtol, atol=atol), np.abs(scores[:, :kk] - s_ref).max()
    full = np.asarray(Q, np.float64) @ np.asarray(D, np.float64).T
    bad = 0
    for b in range(idx.shape[0]):
        for j in range(kk):
            if idx[b, j] != i_ref[b, j]:
                tol = rtol * abs(s_ref[b, j]) + atol
                # the returned doc must score like the oracle's doc at this rank
                assert abs(full[b, idx[b, j]] - s_ref[b, j]) <= tol, (b, j, idx[b, j], i_ref[b, j])
                bad += 1
        assert len(set(idx[b, :kk].tolist())) == kk, "duplicate indices"
    return bad


def model_from_numpy(cfg, sd_np, device, pretrained=True):
    from twotowermlretrieval_b200 import TwoTowerModel
    m = TwoTowerModel(cfg, sd_np["query_encoder.embedding.weight"] if pretrained else None)
    m.load_state_dict({k: torch.tensor(v) for k, v in sd_np.items()})
    return m.to(device)
