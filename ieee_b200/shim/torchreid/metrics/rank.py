"""torchreid/metrics/rank.py of the reference, served by ieee_b200."""
from ieee_b200.metrics.rank import eval_market1501, evaluate_py, evaluate_rank  # noqa: F401
