"""torchreid.metrics surface (upstream torchreid re-exports these two; the IEEE fork's metrics/ has no
__init__.py, so callers there import the submodules -- both spellings work here)."""
from .distance import compute_distance_matrix
from .rank import evaluate_rank

__all__ = ["compute_distance_matrix", "evaluate_rank"]
