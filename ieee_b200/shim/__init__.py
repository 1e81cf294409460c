"""A stand-in ``torchreid`` package holding ONLY the test-time retrieval surface of the IEEE fork, backed by
ieee_b200.  For scripts written against ``torchreid.metrics`` / ``torchreid.utils`` on a
machine where the reference package is not installed:

    PYTHONPATH=$(python -c "import ieee_b200.shim as s; print(s.PATH)") python my_eval_script.py

With the real torchreid installed use ``ieee_b200.patch_torchreid()`` instead; this
shim deliberately has no models, data managers, losses or engines."""
import os

PATH = os.path.dirname(os.path.abspath(__file__))
