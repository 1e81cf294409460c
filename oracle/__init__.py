"""CPU oracle for the IEEE/torchreid retrieval hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it.  The product (``ieee_b200``) never
does, and fails loudly when its CUDA library is missing.

Two layers:

* ``oracle.restatement`` -- a from-scratch NumPy / torch-CPU restatement of the
  reference algorithm (each function cites the reference file:line it follows),
  with a *stable-tie* mode (ties broken by gallery index) that the reference
  leaves undefined (it uses NumPy's unstable argsort).
* ``oracle._ref`` -- the reference's OWN four source files
  (``torchreid/metrics/distance.py``, ``torchreid/metrics/rank.py``,
  ``torchreid/utils/rerank.py``, ``torchreid/metrics/rank_cylib/rank_cy.pyx``)
  compiled, unmodified and from where they lie under ``/root/reference``, into
  extension modules by ``oracle/build_ref.py`` (Cython -> C -> gcc).  Git-ignored,
  but it travels to the GPU box.  Loaded through ``oracle.ref``.

Parity pinning: the reference ships no golden vectors or tests for this path
(SURVEY.md section 4), so the restatement is pinned against (a) ``oracle/_ref``
run live on the same inputs and (b) ``tests/golden/*.npz``, generated in the
build container by ``tests/golden/make_golden.py`` from the reference's Python
sources imported directly from ``/root/reference``.
"""
