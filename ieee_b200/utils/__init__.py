from .reidtools import ranked_lists, visualize_ranked_results
from .rerank import re_ranking

__all__ = ["re_ranking", "visualize_ranked_results", "ranked_lists"]
