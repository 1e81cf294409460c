"""Generate tests/golden/*.npz from the reference's own Python sources.

Run in the BUILD container only (it needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

The reference modules are imported by path, unmodified:
  torchreid/metrics/distance.py, torchreid/metrics/rank.py (with the dead
  ``numpy.lib.function_base`` import pre-seeded), torchreid/utils/rerank.py;
  ``rank_cy`` is the reference's rank_cy.pyx compiled by oracle/build_ref.py.
Inputs are small, seeded and tie-free (so NumPy's unstable argsort has a unique answer) except
for the ``*_ties`` cases, which record the reference run with ``kind='stable'`` forced through
a NumPy proxy (the order the reference leaves undefined).
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference/torchreid"


def load(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    mod = importlib.util.module_from_spec(spec)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        spec.loader.exec_module(mod)
    return mod


class StableNumpy:
    """Proxy handed to a reference module as ``np`` so that argsort is stable; all else is NumPy."""

    def __getattr__(self, item):
        return getattr(np, item)

    @staticmethod
    def argsort(a, axis=-1, kind=None, order=None):
        return np.argsort(a, axis=axis, kind="stable")


def break_ties(d):
    """Nudge equal values in a row apart by single ulps until every row is tie-free."""
    d = d.copy()
    for row in d:
        while True:
            order = np.argsort(row, kind="stable")
            dup = np.nonzero(row[order][1:] == row[order][:-1])[0]
            if dup.size == 0:
                break
            row[order[dup + 1]] = np.nextafter(row[order[dup + 1]], np.float32(np.inf))
    return d


def main():
    sys.modules.setdefault("numpy.lib.function_base", types.SimpleNamespace(_parse_input_dimensions=None))
    dist = load("ref_distance", "metrics/distance.py")
    rank = load("ref_rank", "metrics/rank.py")
    rerank = load("ref_rerank", "utils/rerank.py")
    from oracle import build_ref, ref as compiled
    build_ref.build(verbose=False)
    from ieee_b200.testing import make_retrieval_set

    # ---- distance: distance.py:6-80 -------------------------------------------------------
    g = torch.Generator().manual_seed(11)
    a = torch.relu(torch.randn(24, 96, generator=g))
    b = torch.relu(torch.randn(40, 96, generator=g))
    b[3] = 0.0                                              # all-zero row: cosine distance exactly 1
    np.savez_compressed(os.path.join(HERE, "distance_small.npz"),
                        a=a.numpy(), b=b.numpy(),
                        euclidean=dist.compute_distance_matrix(a, b, "euclidean").numpy(),
                        cosine=dist.compute_distance_matrix(a, b, "cosine").numpy())

    # ---- rank: shapes of rank_cylib/test_cython.py:27-36 (30 x 300, max_rank 5), seeded -----
    rng = np.random.RandomState(7)
    distmat = (rng.rand(30, 300) * 20).astype(np.float32)
    q_pids, g_pids = rng.randint(0, 30, 30), rng.randint(0, 30, 300)
    q_cam, g_cam = rng.randint(0, 5, 30), rng.randint(0, 5, 300)
    assert (np.sort(distmat, 1)[:, 1:] != np.sort(distmat, 1)[:, :-1]).all()
    cmc, mAP = rank.evaluate_rank(distmat, q_pids, g_pids, q_cam, g_cam, max_rank=5)
    cmc_cy, mAP_cy = compiled.evaluate_cy(distmat, q_pids, g_pids, q_cam, g_cam, 5)
    np.savez_compressed(os.path.join(HERE, "rank_cython_shape.npz"), distmat=distmat, q_pids=q_pids,
                        g_pids=g_pids, q_camids=q_cam, g_camids=g_cam, max_rank=5, cmc=cmc, mAP=mAP,
                        cmc_cy=cmc_cy, mAP_cy=mAP_cy)

    # ---- rank: identity-clustered features, some queries without a gallery match -------------
    s = make_retrieval_set(64, 400, 24, 4, dim=128, sigma=2.5, seed=5, distractor_frac=0.2)
    s.q_pids[:3] = 1000                                     # pids absent from the gallery -> invalid queries
    d = break_ties(dist.compute_distance_matrix(s.qf, s.gf).numpy())
    assert (np.sort(d, 1)[:, 1:] != np.sort(d, 1)[:, :-1]).all()
    cmc, mAP = rank.evaluate_rank(d, s.q_pids, s.g_pids, s.q_camids, s.g_camids)  # max_rank default 20
    cmc_cy, mAP_cy = compiled.evaluate_cy(d, s.q_pids, s.g_pids, s.q_camids, s.g_camids, 20)
    np.savez_compressed(os.path.join(HERE, "rank_clustered.npz"), distmat=d, q_pids=s.q_pids, g_pids=s.g_pids,
                        q_camids=s.q_camids, g_camids=s.g_camids, max_rank=20, cmc=cmc, mAP=mAP,
                        cmc_cy=cmc_cy, mAP_cy=mAP_cy)

    # ---- rank with heavy ties: reference run with a stable argsort ----------------------------
    dq = np.round(d / 25.0).astype(np.float32) * 25.0       # quantise -> thousands of equal distances
    rank.np = StableNumpy()
    cmc, mAP = rank.evaluate_rank(dq, s.q_pids, s.g_pids, s.q_camids, s.g_camids, max_rank=20)
    rank.np = np
    np.savez_compressed(os.path.join(HERE, "rank_ties_stable.npz"), distmat=dq, q_pids=s.q_pids, g_pids=s.g_pids,
                        q_camids=s.q_camids, g_camids=s.g_camids, max_rank=20, cmc=cmc, mAP=mAP,
                        n_ties=int((np.sort(dq, 1)[:, 1:] == np.sort(dq, 1)[:, :-1]).sum()))

    # ---- re-ranking: rerank.py:31-113 ---------------------------------------------------------
    s = make_retrieval_set(20, 70, 8, 3, dim=64, sigma=2.0, seed=9)
    qg = dist.compute_distance_matrix(s.qf, s.gf).numpy()
    qq = dist.compute_distance_matrix(s.qf, s.qf).numpy()
    gg = dist.compute_distance_matrix(s.gf, s.gf).numpy()
    rerank.np = StableNumpy()                               # qq/gg are symmetric up to rounding -> make order defined
    out_default = rerank.re_ranking(qg, qq, gg)             # k1=20, k2=6, lambda=0.3
    out_small = rerank.re_ranking(qg, qq, gg, k1=6, k2=3, lambda_value=0.5)
    out_k2_1 = rerank.re_ranking(qg, qq, gg, k1=8, k2=1, lambda_value=0.3)
    rerank.np = np
    out_unstable = rerank.re_ranking(qg, qq, gg)
    np.savez_compressed(os.path.join(HERE, "rerank_small.npz"), qg=qg, qq=qq, gg=gg, out_default=out_default,
                        out_small=out_small, out_k2_1=out_k2_1, out_unstable=out_unstable,
                        q_pids=s.q_pids, g_pids=s.g_pids, q_camids=s.q_camids, g_camids=s.g_camids)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")


if __name__ == "__main__":
    main()
