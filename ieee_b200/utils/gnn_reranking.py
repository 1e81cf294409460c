"""GNN re-ranking (Zhang et al., arXiv 2012.07620) as an alternative ``rerank`` mode -- the algorithm of
torchreid/utils/GPU-Re-Ranking/gnn_reranking.py:27-59 and its two CUDA extensions
(extension/adjacency_matrix/build_adjacency_matrix_kernel.cu:10-17, extension/propagation/gnn_propagate_kernel.cu:8-22).

The reference builds three dense N x N float32 matrices with torch.mm / topk and two scalar scatter / gather kernels,
and hands indices around as floats.  Here the two contractions (X_u X_u^T and A[:Q] A[Q:]^T) run on the tcgen05
kernel (fp32-grade f16x3 arithmetic), the neighbour lists come from the top-k kernel as int32 (ties by index), and
adjacency, symmetrisation, propagation and row normalisation are four small kernels in csrc/rerank.cu.

``gnn_reranking(X_q, X_g, k1, k2)`` keeps the reference's signature and result (the ranked gallery indices, int64
[Q, G], as a NumPy array); ``gnn_reranking_distmat`` returns the negated re-ranked similarity on the device, which
ranks like a distance matrix and feeds ``evaluate_rank`` without ever being sorted.
"""
from __future__ import annotations

import torch

from .. import _lib
from ..metrics.distance import _device_distmat


def gnn_reranking_distmat(X_q: torch.Tensor, X_g: torch.Tensor, k1: int, k2: int) -> torch.Tensor:
    """-(A[:Q] A[Q:]^T): float32 [Q, G] on the device; ascending order = the reference's ranking (:55-57)."""
    _lib.require_cuda()
    dev = X_q.device if X_q.is_cuda else torch.device("cuda", torch.cuda.current_device())
    xq, xg = X_q.to(dev).float(), X_g.to(dev).float()
    Q, G = xq.shape[0], xg.shape[0]
    N = Q + G
    if not 1 <= k2 <= k1 <= min(N, 1024):
        raise ValueError("gnn_reranking needs 1 <= k2 <= k1 <= min(Q + G, 1024), got k1={} k2={}".format(k1, k2))
    lib = _lib.load()
    with torch.cuda.device(dev):
        xu = torch.cat((xq, xg), 0)
        ld = (N + 31) // 32 * 32
        neg = torch.empty((N, ld), dtype=torch.float32, device=dev)[:, :N]
        _device_distmat(xu, xu, "neg_dot", out=neg)                           # -(X_u X_u^T), gnn_reranking.py:31
        A = torch.empty((N, ld), dtype=torch.float32, device=dev)
        ws = torch.empty(lib.ieee_gnn_rerank_workspace_bytes(N, k1), dtype=torch.uint8, device=dev)
        _lib.call("ieee_gnn_rerank", neg.data_ptr(), neg.stride(0), N, k1, k2, A.data_ptr(), A.stride(0), ws.data_ptr(),
                  ws.numel(), _lib.stream())
        del neg
        return _device_distmat(A[:Q, :N], A[Q:, :N], "neg_dot")                # -(cosine similarity), :55


def gnn_reranking(X_q, X_g, k1, k2):
    """Reference signature (gnn_reranking.py:27): returns the ranked gallery indices L, int64 ndarray [Q, G]."""
    d = gnn_reranking_distmat(X_q, X_g, k1, k2)
    return torch.sort(d, dim=1, stable=True)[1].cpu().numpy()                  # presentation only: evaluation never sorts
