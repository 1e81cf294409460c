"""torchreid/utils/rerank.py of the reference, served by ieee_b200."""
from ieee_b200.utils.rerank import re_ranking  # noqa: F401
