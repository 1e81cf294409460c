"""topk_kernel at the Market shape (3368 x 15913, junk-masked): k = 20 and k = 100, CUDA events, L2 flushed / back to back.

    python profiles/r2_topk_probe.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ieee_b200.engine import PackedFeatures, feature_center, packed_distmat
from ieee_b200.metrics.rank import topk_ranked_list
from ieee_b200.testing import market1501_shaped

dev = torch.device("cuda")
s = market1501_shaped(seed=1)
qf, gf = s.qf.to(dev), s.gf.to(dev)
Q, G = qf.shape[0], gf.shape[0]
out = torch.empty((Q, (G + 31) // 32 * 32), device=dev)[:, :G]
lab = [torch.from_numpy(x).to(dev) for x in (s.q_pids, s.q_camids, s.g_pids, s.g_camids)]
c = feature_center(qf)
packed_distmat(PackedFeatures(qf, "euclidean", False, "f16x3", c), PackedFeatures(gf, "euclidean", False, "f16x3", c), out)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for k in (20, 100):
    for _ in range(3):
        topk_ranked_list(out, lab[0], lab[2], lab[1], lab[3], k=k)
    cold, warm = [], []
    for rep in range(16):
        if rep < 8:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); topk_ranked_list(out, lab[0], lab[2], lab[1], lab[3], k=k); b.record(); torch.cuda.synchronize()
        (cold if rep < 8 else warm).append(a.elapsed_time(b) * 1e3)
    print("k = %d: L2 flushed min %.1f us (%.0f GB/s of 4 Q G), back to back min %.1f us"
          % (k, min(cold), 4.0 * Q * G / min(cold) / 1e3, min(warm)), flush=True)
