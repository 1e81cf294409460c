"""CPU restatement of the reference retrieval hot path (NumPy + torch CPU).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): never imported by ``ieee_b200``.

Written from the algorithm, not from the text of the reference; every function
names the reference lines (relative to /root/reference) whose behaviour it
restates.  Pinned against the reference itself (``oracle/_ref`` and
``tests/golden``) by ``tests/test_oracle.py``.

Tie rule.  The reference ranks with ``np.argsort`` (unstable introsort; order of
equal distances is implementation defined -- torchreid/metrics/rank.py:117,
torchreid/utils/rerank.py:48).  ``stable=True`` (default) breaks ties by lower
gallery index first, which is the order the CUDA path implements;
``stable=False`` calls the same unstable argsort as the reference.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------
# distance  (torchreid/metrics/distance.py)
# ----------------------------------------------------------------------------------------------
def compute_distance_matrix(input1: torch.Tensor, input2: torch.Tensor, metric: str = "euclidean") -> torch.Tensor:
    """distance.py:6-46: argument checks and metric dispatch."""
    assert isinstance(input1, torch.Tensor) and isinstance(input2, torch.Tensor)
    assert input1.dim() == 2, "Expected 2-D tensor, but got {}-D".format(input1.dim())
    assert input2.dim() == 2, "Expected 2-D tensor, but got {}-D".format(input2.dim())
    assert input1.size(1) == input2.size(1)
    if metric == "euclidean":
        return euclidean_squared_distance(input1, input2)
    if metric == "cosine":
        return cosine_distance(input1, input2)
    raise ValueError('Unknown distance metric: {}. Please choose either "euclidean" or "cosine"'.format(metric))


def euclidean_squared_distance(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """distance.py:59-64: ||a||^2 (+) ||b||^2 - 2 a.b^T; squared, not clamped, no sqrt.

    The GEMM is issued exactly as the reference issues it (in-place addmm with beta=1,
    alpha=-2 on the broadcast norm sum) so that the CPU BLAS rounding is the same.
    """
    m, n = a.size(0), b.size(0)
    na = (a * a).sum(dim=1, keepdim=True)          # [m,1]
    nb = (b * b).sum(dim=1, keepdim=True).t()      # [1,n]
    out = na.expand(m, n) + nb.expand(m, n)
    out.addmm_(a, b.t(), beta=1, alpha=-2)
    return out


def cosine_distance(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """distance.py:77-80: 1 - normalize(a).normalize(b)^T, norm clamped at 1e-12 (F.normalize eps)."""
    return 1 - torch.mm(F.normalize(a, p=2, dim=1), F.normalize(b, p=2, dim=1).t())


def distance_fp64(a: torch.Tensor, b: torch.Tensor, metric: str = "euclidean") -> torch.Tensor:
    """Same formulas in float64 -- the yardstick both the reference and the GPU path are measured against."""
    a64, b64 = a.double(), b.double()
    if metric == "cosine":
        a64 = a64 / a64.norm(dim=1, keepdim=True).clamp_min(1e-12)
        b64 = b64 / b64.norm(dim=1, keepdim=True).clamp_min(1e-12)
        return 1 - a64 @ b64.t()
    return (a64 * a64).sum(1, keepdim=True) + (b64 * b64).sum(1, keepdim=True).t() - 2 * (a64 @ b64.t())


# ----------------------------------------------------------------------------------------------
# ranking / CMC / mAP  (torchreid/metrics/rank.py, Market-1501 protocol)
# ----------------------------------------------------------------------------------------------
def _argsort_rows(mat: np.ndarray, stable: bool) -> np.ndarray:
    return np.argsort(mat, axis=1, kind="stable") if stable else np.argsort(mat, axis=1)


def eval_market1501(distmat, q_pids, g_pids, q_camids, g_camids, max_rank, stable: bool = True,
                    return_info: bool = False):
    """rank.py:103-171.  Returns (cmc float32[K'], mAP float64) with K' = min(max_rank, G).

    Per query: rank the gallery by ascending distance (:117); drop gallery entries with the
    query's pid AND camid (:136-137); the query is skipped when nothing relevant is left
    (:142-144); CMC row = clipped cumulative hit vector cut at K' (:145-150); AP = mean over the
    relevant items of (hits so far)/(1-based kept rank) in float64 (:155-160).  Averages:
    float32 sum of rows / number of valid queries (:167-168), float64 mean of AP (:169).
    """
    distmat = np.asarray(distmat)
    q_pids, g_pids = np.asarray(q_pids), np.asarray(g_pids)
    q_camids, g_camids = np.asarray(q_camids), np.asarray(g_camids)
    num_q, num_g = distmat.shape
    if num_g < max_rank:
        max_rank = num_g
        print("Note: number of gallery samples is quite small, got {}".format(num_g))
    order = _argsort_rows(distmat, stable)
    rows, aps, first_hit = [], [], np.full(num_q, -1, dtype=np.int64)
    for q in range(num_q):
        o = order[q]
        same_pid = g_pids[o] == q_pids[q]
        junk = same_pid & (g_camids[o] == q_camids[q])
        hits = same_pid[~junk].astype(np.int32)
        if not hits.any():
            continue
        run = hits.cumsum()
        rows.append(np.minimum(run, 1)[:max_rank])
        first_hit[q] = int(np.argmax(hits))
        precision_at_hit = run / (np.arange(hits.size) + 1.0)      # float64
        aps.append((precision_at_hit * hits).sum() / hits.sum())
    assert len(aps) > 0, "Error: all query identities do not appear in gallery"
    cmc = np.asarray(rows).astype(np.float32).sum(0) / float(len(aps))
    mAP = np.mean(aps)
    if return_info:
        return cmc, mAP, {"num_valid": len(aps), "first_hit": first_hit, "ap": np.asarray(aps)}
    return cmc, mAP


def evaluate_rank(distmat, q_pids, g_pids, q_camids, g_camids, max_rank=20, use_metric_cuhk03=False,
                  use_cython=True, stable: bool = True):
    """rank.py:246-287: in this fork always the Python Market-1501 protocol; ``use_cython`` is ignored (:278-287)."""
    if use_metric_cuhk03:
        # rank.py:236-239 calls the 8-argument eval_cuhk03 with 6 arguments -> TypeError in the reference.
        raise TypeError("eval_cuhk03() missing 2 required positional arguments")
    return eval_market1501(distmat, q_pids, g_pids, q_camids, g_camids, max_rank, stable=stable)


def eval_cuhk03(distmat, q_pids, g_pids, q_camids, g_camids, max_rank, q_timeids=None, g_timeids=None,
                num_repeats: int = 10, stable: bool = True):
    """Single-gallery-shot protocol: rank.py:24-100 (this fork: 8 arguments, time ids join the junk rule, :48) and
    rank_cy.pyx:37-153 (the upstream 6-argument form, the one that still runs: rank.py's uses np.bool, gone from NumPy).

    Per query: rank; drop gallery items with the query's pid AND camera (AND time id when given); skip the query when
    nothing relevant is left; then `num_repeats` times keep ONE random item per gallery identity -- identities visited in
    order of first appearance in the kept ranked list, one ``np.random.choice(positions)`` each (rank.py:66-72) -- and
    accumulate the clipped cumulative hit vector of that sample; AP comes from the unsampled kept list (:81-86).
    RNG parity = NumPy's global generator consumed in exactly that order: seed it, call, compare.
    Returns (cmc float32[K'], mAP float64)."""
    distmat = np.asarray(distmat)
    q_pids, g_pids, q_camids, g_camids = (np.asarray(a) for a in (q_pids, g_pids, q_camids, g_camids))
    num_q, num_g = distmat.shape
    if num_g < max_rank:
        max_rank = num_g
        print("Note: number of gallery samples is quite small, got {}".format(num_g))
    order = _argsort_rows(distmat, stable)
    rows, aps = [], []
    for q in range(num_q):
        o = order[q]
        junk = (g_pids[o] == q_pids[q]) & (g_camids[o] == q_camids[q])
        if q_timeids is not None:
            junk &= np.asarray(g_timeids)[o] == np.asarray(q_timeids)[q]
        hits = (g_pids[o] == q_pids[q])[~junk].astype(np.int32)
        if not hits.any():
            continue
        kept_pids = g_pids[o][~junk]
        groups = {}
        for pos, pid in enumerate(kept_pids.tolist()):
            groups.setdefault(pid, []).append(pos)
        acc = np.zeros(max_rank, dtype=np.float32)
        for _ in range(num_repeats):
            pick = np.zeros(hits.size, dtype=bool)
            for positions in groups.values():
                pick[np.random.choice(positions)] = True
            run = np.minimum(hits[pick].cumsum(), 1)[:max_rank].astype(np.float32)
            acc[: run.size] += run                               # (fewer identities than max_rank: rank.py would fail to add)
        rows.append(acc / num_repeats)
        run = hits.cumsum()
        aps.append(float((run / (np.arange(hits.size) + 1.0) * hits).sum() / hits.sum()))
    assert len(rows) > 0, "Error: all query identities do not appear in gallery"
    cmc = np.asarray(rows).astype(np.float32).sum(0) / float(len(rows))
    return cmc.astype(np.float32), float(np.mean(aps))


def kept_positions(distmat, q_pids, g_pids, q_camids, g_camids):
    """Sort-free form of rank.py:136-160, used to check kernel intermediates.

    For query q and each relevant gallery item r (same pid, other camera) returns the 0-based
    position of r among the kept items: #{kept g : (d[q,g], g) <lex (d[q,r], r)}.
    Returns a list (per query) of sorted int64 arrays (empty -> invalid query).
    """
    distmat = np.asarray(distmat)
    out = []
    idx = np.arange(distmat.shape[1])
    for q in range(distmat.shape[0]):
        same = np.asarray(g_pids) == q_pids[q]
        junk = same & (np.asarray(g_camids) == q_camids[q])
        rel = np.nonzero(same & ~junk)[0]
        d = distmat[q]
        kd, ki = d[~junk], idx[~junk]
        pos = []
        for r in rel:
            with np.errstate(invalid="ignore"):
                less = (kd < d[r]) | ((kd == d[r]) & (ki < r))
                if np.isnan(d[r]):     # NaN sorts last (NumPy); NaNs tie with each other
                    less = ~np.isnan(kd) | (np.isnan(kd) & (ki < r))
            pos.append(int(less.sum()))
        out.append(np.sort(np.asarray(pos, dtype=np.int64)))
    return out


def metrics_from_positions(positions, max_rank, num_g):
    """CMC / mAP from kept positions: cmc_row[j] = [j >= p_1], AP = mean_k k/(p_k+1)."""
    max_rank = min(max_rank, num_g)
    hits = np.zeros(max_rank, dtype=np.int64)
    aps = []
    for p in positions:
        if p.size == 0:
            continue
        hits[int(p[0]):] += 1          # empty slice when the first hit is beyond max_rank
        aps.append(float(((np.arange(p.size) + 1.0) / (p + 1.0)).sum() / p.size))
    cmc = hits.astype(np.float32) / np.float32(len(aps))
    return cmc, float(np.mean(aps)), len(aps)


def mean_inverse_negative_penalty(distmat, q_pids, g_pids, q_camids, g_camids, stable: bool = True) -> float:
    """mINP.  The reference only NAMES this metric (README.rst:45, pointing at Ye et al., TPAMI 2021 and
    github.com/mangye16/ReID-Survey); it has no code for it, so this restates the published definition on top of the
    reference's own ranked list (rank.py:117,136-144): for every valid query, INP = R / (1-based position of the
    hardest, i.e. last, relevant item among the kept ones); mINP = mean over valid queries.  Parity unpinned by any
    reference vector (there is none); pinned only to kept_positions() below by tests/test_oracle.py."""
    distmat = np.asarray(distmat)
    order = _argsort_rows(distmat, stable)
    g_pids, g_camids = np.asarray(g_pids), np.asarray(g_camids)
    inps = []
    for q in range(distmat.shape[0]):
        o = order[q]
        keep = ~((g_pids[o] == q_pids[q]) & (g_camids[o] == q_camids[q]))
        raw = (g_pids[o] == q_pids[q])[keep]
        if not raw.any():
            continue
        hard = int(np.nonzero(raw)[0].max())
        inps.append(float(raw.sum()) / (hard + 1.0))
    return float(np.mean(inps)) if inps else 0.0


def topk_kept(distmat, q_pids, g_pids, q_camids, g_camids, k, stable: bool = True):
    """The first k entries of each query's junk-filtered ranked list (rank.py:117,136-140;
    what torchreid/utils/reidtools.py:49,111 walks).  Returns (idx int64 [Q,k], dist [Q,k]);
    rows with fewer than k kept items are padded with -1 / +inf."""
    distmat = np.asarray(distmat)
    Q, G = distmat.shape
    order = _argsort_rows(distmat, stable)
    idx = np.full((Q, k), -1, dtype=np.int64)
    val = np.full((Q, k), np.inf, dtype=distmat.dtype)
    for q in range(Q):
        o = order[q]
        junk = (np.asarray(g_pids)[o] == q_pids[q]) & (np.asarray(g_camids)[o] == q_camids[q])
        kept = o[~junk][:k]
        idx[q, :kept.size] = kept
        val[q, :kept.size] = distmat[q, kept]
    return idx, val


def count_row_ties(distmat) -> int:
    """Number of adjacent equal pairs in each sorted row (0 -> the unstable reference order is unique)."""
    s = np.sort(np.asarray(distmat), axis=1)
    return int((s[:, 1:] == s[:, :-1]).sum())


# ----------------------------------------------------------------------------------------------
# k-reciprocal re-ranking  (torchreid/utils/rerank.py)
# ----------------------------------------------------------------------------------------------
def rerank_original_dist(q_g_dist, q_q_dist, g_g_dist) -> np.ndarray:
    """rerank.py:36-46: N x N block matrix [[qq, qg], [qg^T, gg]], squared element-wise (again),
    as float32, each column divided by its maximum, then transposed."""
    top = np.concatenate([q_q_dist, q_g_dist], axis=1)
    bot = np.concatenate([q_g_dist.T, g_g_dist], axis=1)
    m = np.power(np.concatenate([top, bot], axis=0), 2).astype(np.float32)
    return np.transpose(1.0 * m / np.max(m, axis=0))


def _k_reciprocal(rank: np.ndarray, i: int, k: int) -> np.ndarray:
    """rerank.py:56-59 / :63-71: members j of i's first k+1 neighbours whose own first k+1 contain i."""
    fwd = rank[i, : k + 1]
    back = rank[fwd, : k + 1]
    return fwd[np.where(back == i)[0]]


def re_ranking(q_g_dist, q_q_dist, g_g_dist, k1=20, k2=6, lambda_value=0.3, stable: bool = True,
               return_parts: bool = False):
    """rerank.py:31-113, restated with a row-sparse V (same floating-point operation order).

    Stages: normalised N x N distances; neighbour ranking; k-reciprocal sets
    with the half-k expansion rule (strict > 2/3 overlap against the un-expanded set, :72-75);
    Gaussian weights normalised over the sorted unique set (:80-82); k2 query expansion as the
    mean of k2 full rows (:84-89); Jaccard distance through the inverted index, accumulated in
    ascending column order (:91-106); lambda blend and the [:Q, Q:] slice (:108-113).
    """
    q_g_dist, q_q_dist, g_g_dist = map(np.asarray, (q_g_dist, q_q_dist, g_g_dist))
    orig = rerank_original_dist(q_g_dist, q_q_dist, g_g_dist)
    Q = q_g_dist.shape[0]
    N = Q + q_g_dist.shape[1]
    rank = _argsort_rows(orig, stable).astype(np.int32)
    k_half = int(np.around(k1 / 2.0))

    V = np.zeros((N, N), dtype=np.float32)
    for i in range(N):
        base = _k_reciprocal(rank, i, k1)
        grown = base
        for cand in base:
            cset = _k_reciprocal(rank, int(cand), k_half)
            if len(np.intersect1d(cset, base)) > 2.0 / 3 * len(cset):
                grown = np.append(grown, cset)
        members = np.unique(grown)
        w = np.exp(-orig[i, members])
        V[i, members] = 1.0 * w / np.sum(w)
    if k2 != 1:
        V_qe = np.zeros_like(V, dtype=np.float32)
        for i in range(N):
            V_qe[i, :] = np.mean(V[rank[i, :k2], :], axis=0)
        V = V_qe
    inv_index = [np.where(V[:, c] != 0)[0] for c in range(N)]
    jaccard = np.zeros((Q, N), dtype=np.float32)
    for i in range(Q):
        acc = np.zeros(N, dtype=np.float32)
        for c in np.where(V[i, :] != 0)[0]:
            rows = inv_index[c]
            acc[rows] = acc[rows] + np.minimum(V[i, c], V[rows, c])
        jaccard[i] = 1 - acc / (2.0 - acc)
    final = jaccard * (1 - lambda_value) + orig[:Q] * lambda_value
    out = final[:, Q:]
    if return_parts:
        return out, {"orig": orig, "rank": rank, "V": V, "jaccard": jaccard}
    return out


# ----------------------------------------------------------------------------------------------
# GNN re-ranking  (torchreid/utils/GPU-Re-Ranking/gnn_reranking.py + its two CUDA extensions)
# PARITY UNPINNED: the reference needs its extensions compiled and a CUDA device, neither of which the build container
# has, and it ships no vectors; this restatement follows the source line by line and is what the GPU tests check.
# ----------------------------------------------------------------------------------------------
def gnn_reranking(X_q, X_g, k1, k2, return_similarity: bool = False):
    """gnn_reranking.py:27-59.  Neighbour ties go to the lower index (torch.topk leaves them undefined)."""
    X_q, X_g = np.asarray(X_q, dtype=np.float32), np.asarray(X_g, dtype=np.float32)
    Q = X_q.shape[0]
    X_u = np.concatenate((X_q, X_g), 0)
    score = X_u @ X_u.T                                            # :31
    rank = np.argsort(-score, axis=1, kind="stable")[:, :k1]       # :36-38 (largest first, sorted)
    S = np.take_along_axis(score, rank, 1)
    N = score.shape[0]
    A = np.zeros((N, N), dtype=np.float32)                          # build_adjacency_matrix_kernel.cu:10-17
    A[np.arange(N)[:, None], rank] = 1.0
    S = S * S                                                       # :42
    if k2 != 1:
        for _ in range(2):                                          # :45-53
            A = A + A.T
            out = np.zeros_like(A)
            for j in range(k2):                                     # gnn_propagate_kernel.cu:13-20, summed in j order
                out += A[rank[:, j]] * S[:, j:j + 1]
            A = out / np.sqrt((out.astype(np.float64) ** 2).sum(1, keepdims=True)).astype(np.float32)
    cos = A[:Q] @ A[Q:].T                                           # :55
    if return_similarity:
        return cos
    return np.argsort(-cos, axis=1, kind="stable")                  # :58-59
