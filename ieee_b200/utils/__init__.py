from .rerank import re_ranking

__all__ = ["re_ranking"]
