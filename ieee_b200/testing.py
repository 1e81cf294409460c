"""Synthetic workloads shaped like the reference's evaluation sets.

Features imitate ``IEEE3modalPart``'s eval output (torchreid/models/ieee3modalPart.py:497-505):
``D = 2304 = 3 modalities x 6 parts x 128``, float32, post-ReLU (non-negative, roughly half zeros),
clustered by identity so that mAP lands well inside (0, 1).  Everything is generated on the CPU
with a seeded ``torch.Generator`` so a seed names the same workload on every machine.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

FEATURE_DIM = 2304  # ieee3modalPart.py:298-303,497-502


@dataclass
class RetrievalSet:
    qf: torch.Tensor        # [Q, D] float32
    gf: torch.Tensor        # [G, D] float32
    q_pids: np.ndarray      # [Q] int64
    g_pids: np.ndarray      # [G] int64
    q_camids: np.ndarray    # [Q] int64
    g_camids: np.ndarray    # [G] int64
    name: str = ""


def clustered_features(pids: torch.Tensor, centers: torch.Tensor, sigma: float, gen: torch.Generator,
                       chunk: int = 8192) -> torch.Tensor:
    out = torch.empty(pids.numel(), centers.shape[1], dtype=torch.float32)
    for s in range(0, pids.numel(), chunk):
        e = min(s + chunk, pids.numel())
        noise = torch.randn(e - s, centers.shape[1], generator=gen)
        out[s:e] = torch.relu(centers[pids[s:e]] + sigma * noise)
    return out


def make_retrieval_set(num_q: int, num_g: int, num_pids: int, num_cams: int, dim: int = FEATURE_DIM,
                       sigma: float = 3.5, seed: int = 0, distractor_frac: float = 0.0,
                       same_set: bool = False, name: str = "") -> RetrievalSet:
    """Identity-clustered query/gallery sets.

    ``same_set=True`` reproduces RGBNT201, where the query and gallery loaders read the same
    directory (torchreid/data/datasets/image/RGBNT201.py:33-34,42-43): gallery == query.
    ``distractor_frac`` adds Market-1501-style pid-0 distractors that no query matches.
    """
    gen = torch.Generator().manual_seed(seed)
    centers = torch.randn(num_pids + 1, dim, generator=gen)
    q_pids = torch.randint(1, num_pids + 1, (num_q,), generator=gen)
    q_cams = torch.randint(0, num_cams, (num_q,), generator=gen)
    qf = clustered_features(q_pids, centers, sigma, gen)
    if same_set:
        g_pids, g_cams, gf = q_pids.clone(), q_cams.clone(), qf.clone()
    else:
        g_pids = torch.randint(1, num_pids + 1, (num_g,), generator=gen)
        if distractor_frac > 0:
            mask = torch.rand(num_g, generator=gen) < distractor_frac
            g_pids[mask] = 0
        g_cams = torch.randint(0, num_cams, (num_g,), generator=gen)
        gf = clustered_features(g_pids, centers, sigma, gen)
    return RetrievalSet(qf, gf, q_pids.numpy().astype(np.int64), g_pids.numpy().astype(np.int64),
                        q_cams.numpy().astype(np.int64), g_cams.numpy().astype(np.int64), name)


def rgbnt201_shaped(seed: int = 0, sigma: float = 3.5) -> RetrievalSet:
    """Config C1: Q = G = 836 (query set == gallery set), 30 ids, 4 cameras."""
    return make_retrieval_set(836, 836, 30, 4, sigma=sigma, seed=seed, same_set=True, name="rgbnt201_shaped")


def market1501_shaped(seed: int = 1, sigma: float = 2.75, num_q: int = 3368) -> RetrievalSet:
    """Config C2: 3368 queries x 15913 gallery (torchreid/data/datasets/image/market1501.py:21), 751 ids, 6 cameras."""
    return make_retrieval_set(num_q, 15913, 751, 6, sigma=sigma, seed=seed, distractor_frac=0.17,
                              name="market1501_shaped")
