"""Loader for ``oracle/_ref`` -- the reference's own sources, compiled (see build_ref.py).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Exposes the reference entry points under their own names:

    ref.compute_distance_matrix   torchreid/metrics/distance.py:6
    ref.evaluate_rank             torchreid/metrics/rank.py:246   (always the Python Market-1501 protocol, rank.py:284)
    ref.eval_market1501           torchreid/metrics/rank.py:103
    ref.evaluate_cy               torchreid/metrics/rank_cylib/rank_cy.pyx:26
    ref.re_ranking                torchreid/utils/rerank.py:31
"""
from __future__ import annotations

import importlib.machinery
import importlib.util
import sys
import types
import warnings

from . import build_ref

_cache: dict = {}


def available() -> bool:
    return build_ref.built()


def _load(name: str):
    if name in _cache:
        return _cache[name]
    path = build_ref.ext_path(name)
    if name == "rank":
        # rank.py:8 imports a private NumPy symbol that NumPy >= 2 no longer has (dead import,
        # never used).  Pre-seed it so the unmodified module initialises.
        sys.modules.setdefault("numpy.lib.function_base",
                               types.SimpleNamespace(_parse_input_dimensions=None))
    loader = importlib.machinery.ExtensionFileLoader(name, path)
    spec = importlib.util.spec_from_file_location(name, path, loader=loader)
    mod = importlib.util.module_from_spec(spec)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        loader.exec_module(mod)
    _cache[name] = mod
    return mod


def compute_distance_matrix(input1, input2, metric="euclidean"):
    return _load("distance").compute_distance_matrix(input1, input2, metric)


def evaluate_rank(distmat, q_pids, g_pids, q_camids, g_camids, max_rank=20,
                  use_metric_cuhk03=False, use_cython=True):
    return _load("rank").evaluate_rank(distmat, q_pids, g_pids, q_camids, g_camids, max_rank,
                                       use_metric_cuhk03, use_cython)


def eval_market1501(distmat, q_pids, g_pids, q_camids, g_camids, max_rank):
    return _load("rank").eval_market1501(distmat, q_pids, g_pids, q_camids, g_camids, max_rank)


def evaluate_cy(distmat, q_pids, g_pids, q_camids, g_camids, max_rank, use_metric_cuhk03=False):
    return _load("rank_cy").evaluate_cy(distmat, q_pids, g_pids, q_camids, g_camids, max_rank,
                                        use_metric_cuhk03)


def re_ranking(q_g_dist, q_q_dist, g_g_dist, k1=20, k2=6, lambda_value=0.3):
    return _load("rerank").re_ranking(q_g_dist, q_q_dist, g_g_dist, k1, k2, lambda_value)
