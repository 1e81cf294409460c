"""torchreid surface of the IEEE fork's test-time retrieval path, served by ieee_b200 (see ieee_b200/shim)."""
from . import metrics, utils  # noqa: F401

__version__ = "1.4.0+ieee_b200"
