"""torchreid.metrics: upstream torchreid re-exports the two functions (the fork has no metrics/__init__.py and
imports the submodules, engine.py:18-19) -- both spellings resolve."""
from ieee_b200.metrics.distance import compute_distance_matrix
from ieee_b200.metrics.rank import evaluate_rank

from . import distance, rank  # noqa: F401

__all__ = ["compute_distance_matrix", "evaluate_rank"]
