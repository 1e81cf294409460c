"""ieee_b200 -- B200 (sm_100a) implementation of the test-time retrieval hot path of ziwang1121/IEEE.

Drop-in surface (same names / arguments as the torchreid fork):

    ieee_b200.metrics.compute_distance_matrix   torchreid/metrics/distance.py:6
    ieee_b200.metrics.evaluate_rank             torchreid/metrics/rank.py:246
    ieee_b200.utils.re_ranking                  torchreid/utils/rerank.py:31
    ieee_b200.utils.visualize_ranked_results    torchreid/utils/reidtools.py:18
    ieee_b200.engine.evaluate                   torchreid/engine/engine.py:391-417 (device-resident)

All arithmetic runs in ``libieee_b200.so`` (hand-written CUDA behind the C ABI of include/ieee_b200.h);
there is no CPU fallback.
"""
from . import metrics, utils  # noqa: F401

__version__ = "0.1.0"


def patch_torchreid():
    """Route an imported torchreid through this package (the fork imports the submodules directly:
    torchreid/engine/engine.py:18-19 and torchreid/utils/__init__.py:4)."""
    import importlib
    import sys

    from .metrics import distance as _d, rank as _r
    from .utils import reidtools as _rt, rerank as _rr

    for name, mod in (("torchreid.metrics.distance", _d), ("torchreid.metrics.rank", _r),
                      ("torchreid.utils.rerank", _rr), ("torchreid.utils.reidtools", _rt)):
        sys.modules[name] = mod
    eng = sys.modules.get("torchreid.engine.engine")
    if eng is not None:
        eng.compute_distance_matrix = _d.compute_distance_matrix
        eng.evaluate_rank = _r.evaluate_rank
        eng.re_ranking = _rr.re_ranking
        eng.visualize_ranked_results = _rt.visualize_ranked_results      # engine.py:26
    tu = sys.modules.get("torchreid.utils")
    if tu is not None:
        tu.re_ranking = _rr.re_ranking
        tu.visualize_ranked_results = _rt.visualize_ranked_results
    importlib.invalidate_caches()
