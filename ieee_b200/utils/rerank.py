"""Drop-in for torchreid/utils/rerank.py (k-reciprocal re-ranking, Zhong et al. CVPR 2017)."""
from __future__ import annotations

import numpy as np
import torch

from .. import _lib

__all__ = ["re_ranking"]


def re_ranking_device(q_g: torch.Tensor, q_q: torch.Tensor, g_g: torch.Tensor, k1=20, k2=6, lambda_value=0.3):
    """CUDA float32 matrices in, CUDA float32 [Q, G] out."""
    Q, G = q_g.shape
    assert q_q.shape == (Q, Q) and g_g.shape == (G, G)
    dev = q_g.device
    out = torch.empty((Q, G), dtype=torch.float32, device=dev)
    lib = _lib.load()
    ws_bytes = lib.ieee_rerank_workspace_bytes(Q, G, k1, k2)
    ws = torch.empty(max(ws_bytes, 256), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.call("ieee_rerank", q_g.data_ptr(), q_g.stride(0), q_q.data_ptr(), q_q.stride(0), g_g.data_ptr(),
                  g_g.stride(0), Q, G, k1, k2, float(lambda_value), out.data_ptr(), out.stride(0), ws.data_ptr(),
                  ws_bytes, _lib.stream())
    return out


def re_ranking(q_g_dist, q_q_dist, g_g_dist, k1=20, k2=6, lambda_value=0.3):
    """Reference: rerank.py:31.  NumPy (or torch) distance matrices in, float32 ndarray [Q, G] out
    (a CUDA tensor in gives a CUDA tensor out)."""
    _lib.require_cuda()
    on_device = isinstance(q_g_dist, torch.Tensor) and q_g_dist.is_cuda
    dev = q_g_dist.device if on_device else torch.device("cuda", torch.cuda.current_device())

    def dev32(x):
        t = x if isinstance(x, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(x))
        t = t.to(device=dev, dtype=torch.float32, non_blocking=True)
        return t if t.stride(-1) == 1 else t.contiguous()

    out = re_ranking_device(dev32(q_g_dist), dev32(q_q_dist), dev32(g_g_dist), k1, k2, lambda_value)
    return out if on_device else out.cpu().numpy()
