"""GPU parity: k-reciprocal re-ranking (torchreid/utils/rerank.py:31-113) vs golden vectors and the CPU oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import ref, restatement as R
from ieee_b200.metrics import compute_distance_matrix, evaluate_rank
from ieee_b200.utils import re_ranking
from ieee_b200.testing import make_retrieval_set, rgbnt201_shaped

pytestmark = pytest.mark.gpu

ATOL = 1e-5   # fp32 exp / normalisation differ in the last bits between NumPy and CUDA


@pytest.mark.parametrize("key,kw", [("out_default", {}), ("out_small", dict(k1=6, k2=3, lambda_value=0.5)),
                                    ("out_k2_1", dict(k1=8, k2=1, lambda_value=0.3))])
def test_golden_rerank(golden_dir, key, kw):
    g = np.load(os.path.join(golden_dir, "rerank_small.npz"))
    out = re_ranking(g["qg"], g["qq"], g["gg"], **kw)
    assert isinstance(out, np.ndarray) and out.dtype == np.float32 and out.shape == g[key].shape
    np.testing.assert_allclose(out, g[key], rtol=0, atol=ATOL)


def intermediate_checks(qg, qq, gg, k1, k2, lam):
    out_o, parts = R.re_ranking(qg, qq, gg, k1=k1, k2=k2, lambda_value=lam, return_parts=True)
    out = re_ranking(torch.from_numpy(qg).cuda(), torch.from_numpy(qq).cuda(), torch.from_numpy(gg).cuda(), k1, k2, lam)
    assert out.is_cuda
    out = out.cpu().numpy()
    err = np.abs(out - out_o)
    assert err.max() <= ATOL, f"max abs err {err.max():.3e}"
    # ranking agreement: the re-ranked order of the first 10 gallery items per query
    top_o = np.argsort(out_o, axis=1, kind="stable")[:, :10]
    top = np.argsort(out, axis=1, kind="stable")[:, :10]
    agree = (top == top_o).mean()
    assert agree > 0.98, agree
    return out, out_o


def test_rgbnt201_shaped_rerank_config3():
    """Config C3: k1=20, k2=6, lambda=0.3 on the RGBNT201-shaped set (N = 1672, query set == gallery set)."""
    s = rgbnt201_shaped()
    qg = R.compute_distance_matrix(s.qf, s.gf).numpy()
    qq = R.compute_distance_matrix(s.qf, s.qf).numpy()
    gg = R.compute_distance_matrix(s.gf, s.gf).numpy()
    out, out_o = intermediate_checks(qg, qq, gg, 20, 6, 0.3)
    cmc, mAP = evaluate_rank(out, s.q_pids, s.g_pids, s.q_camids, s.g_camids)
    cmc_o, map_o = R.evaluate_rank(out_o, s.q_pids, s.g_pids, s.q_camids, s.g_camids)
    assert abs(mAP - map_o) < 2e-3 and np.abs(cmc - cmc_o).max() < 5e-3
    cmc_plain, map_plain = R.evaluate_rank(qg, s.q_pids, s.g_pids, s.q_camids, s.g_camids)
    assert mAP > map_plain                      # re-ranking helps on clustered identities


@pytest.mark.parametrize("shape,k1,k2,lam", [((37, 150), 20, 6, 0.3), ((50, 333), 10, 4, 0.1), ((8, 40), 5, 1, 0.7),
                                             ((120, 600), 30, 8, 0.3)])
def test_rerank_shapes(shape, k1, k2, lam):
    s = make_retrieval_set(shape[0], shape[1], 12, 3, dim=96, sigma=2.0, seed=sum(shape))
    qg = R.compute_distance_matrix(s.qf, s.gf).numpy()
    qq = R.compute_distance_matrix(s.qf, s.qf).numpy()
    gg = R.compute_distance_matrix(s.gf, s.gf).numpy()
    intermediate_checks(qg, qq, gg, k1, k2, lam)


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")
def test_rerank_against_compiled_reference():
    """The unmodified reference (unstable argsort) on a tie-free case."""
    s = make_retrieval_set(40, 200, 10, 3, dim=64, sigma=2.0, seed=77)
    qg = ref.compute_distance_matrix(s.qf, s.gf).numpy()
    qq = ref.compute_distance_matrix(s.qf, s.qf).numpy()
    gg = ref.compute_distance_matrix(s.gf, s.gf).numpy()
    theirs = ref.re_ranking(qg, qq, gg)
    ours = re_ranking(qg, qq, gg)
    orig = R.rerank_original_dist(qg, qq, gg)
    if R.count_row_ties(orig[:, :]) == 0:
        np.testing.assert_allclose(ours, theirs, rtol=0, atol=ATOL)
    else:   # ties among the first k1+1 neighbours are what matters; compare with the stable oracle instead
        np.testing.assert_allclose(ours, R.re_ranking(qg, qq, gg), rtol=0, atol=ATOL)


def test_engine_evaluate_with_rerank(capsys):
    from ieee_b200.engine import evaluate
    s = make_retrieval_set(100, 500, 15, 3, dim=256, sigma=2.5, seed=5)
    cmc, mAP = evaluate(s.qf, s.gf, s.q_pids, s.g_pids, s.q_camids, s.g_camids, rerank=True, ranks=[1, 5, 10])
    text = capsys.readouterr().out
    assert "Applying person re-ranking" in text and "mAP: " in text and "Rank-1  :" in text      # engine.py:402,420-425
    qg = R.compute_distance_matrix(s.qf, s.gf).numpy()
    qq = R.compute_distance_matrix(s.qf, s.qf).numpy()
    gg = R.compute_distance_matrix(s.gf, s.gf).numpy()
    cmc_o, map_o = R.evaluate_rank(R.re_ranking(qg, qq, gg), s.q_pids, s.g_pids, s.q_camids, s.g_camids)
    assert abs(mAP - map_o) < 5e-3


@pytest.mark.parametrize("k1,k2", [(26, 7), (10, 1), (20, 6)])
def test_gnn_reranking_matches_restatement(k1, k2):
    """GNN re-ranking (gnn_reranking.py:27-59) against the oracle restatement on L2-normalised features, as the
    reference's own driver feeds it: re-ranked similarity within 1e-5, ranked lists equal except where the
    restatement's similarities of the swapped candidates are within 1e-4 of each other."""
    from ieee_b200.utils.gnn_reranking import gnn_reranking, gnn_reranking_distmat
    s = make_retrieval_set(60, 300, 12, 3, dim=256, sigma=1.5, seed=k1)
    xq = torch.nn.functional.normalize(s.qf, dim=1)
    xg = torch.nn.functional.normalize(s.gf, dim=1)
    cos_o = R.gnn_reranking(xq.numpy(), xg.numpy(), k1, k2, return_similarity=True)
    d = gnn_reranking_distmat(xq.cuda(), xg.cuda(), k1, k2).cpu().numpy()
    assert d.shape == (60, 300) and np.abs(-d - cos_o).max() < 1e-5
    L = gnn_reranking(xq, xg, k1, k2)
    L_o = R.gnn_reranking(xq.numpy(), xg.numpy(), k1, k2)
    assert L.shape == L_o.shape == (60, 300) and L.dtype == np.int64
    differ = L != L_o                                     # exact ties (the many zeros) go by index on both sides
    a, b = np.take_along_axis(cos_o, L, 1), np.take_along_axis(cos_o, L_o, 1)
    assert differ.mean() < 0.05 and np.abs(a - b)[differ].max(initial=0.0) < 1e-4   # only near-ties may swap
    # as a rerank mode: the negated similarity ranks like a distance matrix
    cmc, mAP = evaluate_rank(d, s.q_pids, s.g_pids, s.q_camids, s.g_camids)
    cmc_o, map_o = R.evaluate_rank(d, s.q_pids, s.g_pids, s.q_camids, s.g_camids)
    assert np.array_equal(cmc, cmc_o) and abs(mAP - map_o) < 1e-9
    with pytest.raises(ValueError):
        gnn_reranking_distmat(xq.cuda(), xg.cuda(), 5, 6)
