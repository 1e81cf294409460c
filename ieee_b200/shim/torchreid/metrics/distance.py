"""torchreid/metrics/distance.py of the reference, served by ieee_b200."""
from ieee_b200.metrics.distance import compute_distance_matrix, cosine_distance, euclidean_squared_distance  # noqa: F401
