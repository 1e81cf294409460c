"""torchreid/utils/reidtools.py of the reference, served by ieee_b200."""
from ieee_b200.utils.reidtools import visualize_ranked_results  # noqa: F401
