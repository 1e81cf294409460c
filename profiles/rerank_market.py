"""k-reciprocal re-ranking at Market-1501 scale (Q = 3368, G = 15913, N = 19281): timing of ieee_rerank and its effect on
the metrics.  (The reference's NumPy implementation needs minutes here; parity is pinned at smaller sizes.)"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ieee_b200.metrics.distance import _device_distmat
from ieee_b200.metrics.rank import evaluate_device
from ieee_b200.testing import market1501_shaped
from ieee_b200.utils.rerank import re_ranking_device

s = market1501_shaped()
dev = torch.device("cuda")
qf, gf = s.qf.to(dev), s.gf.to(dev)
t0 = time.perf_counter()
qg, qq, gg = _device_distmat(qf, gf, "euclidean"), _device_distmat(qf, qf, "euclidean"), _device_distmat(gf, gf, "euclidean")
torch.cuda.synchronize()
t1 = time.perf_counter()
out = re_ranking_device(qg, qq, gg)
torch.cuda.synchronize()
t2 = time.perf_counter()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
out = re_ranking_device(qg, qq, gg)
b.record()
torch.cuda.synchronize()
cmc0, s0, _ = evaluate_device(qg, s.q_pids, s.g_pids, s.q_camids, s.g_camids, 20)
cmc1, s1, _ = evaluate_device(out, s.q_pids, s.g_pids, s.q_camids, s.g_camids, 20)
print("three distance matrices (qg, qq, gg): %.1f ms; re_ranking first call %.1f ms, steady %.1f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, a.elapsed_time(b)))
print("mAP %.4f -> %.4f, rank-1 %.4f -> %.4f" % (s0.mAP, s1.mAP, cmc0[0].item(), cmc1[0].item()))
