"""twotowermlretrieval_b200 — the data-parallel hot path of jpe17/TwoTowerMLRetrieval
(GRU two-tower encode, cosine triplet step, exact cosine top-k, TF-IDF hybrid rerank) as
hand-written sm_100a CUDA behind the reference's Python API.  See DESIGN.md."""
from .model import ModelFactory, RNNEncoder, TwoTowerModel, triplet_loss_cosine  # noqa: F401

__all__ = ["RNNEncoder", "TwoTowerModel", "triplet_loss_cosine", "ModelFactory"]
