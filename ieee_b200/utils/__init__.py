from .reidtools import ranked_lists, visualize_ranked_results
from .rerank import re_ranking
from .gnn_reranking import gnn_reranking

__all__ = ["re_ranking", "gnn_reranking", "visualize_ranked_results", "ranked_lists"]
