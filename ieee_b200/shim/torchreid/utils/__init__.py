"""torchreid.utils: the two names on the retrieval path (utils/__init__.py:4,7)."""
from ieee_b200.utils.reidtools import visualize_ranked_results
from ieee_b200.utils.rerank import re_ranking

from . import reidtools, rerank  # noqa: F401

__all__ = ["re_ranking", "visualize_ranked_results"]
