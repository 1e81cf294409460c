"""rank_count as rank 0 of an N-way gallery-sharded evaluation sees it, emulated on ONE GPU (bench workload:
Q = 3368 * N queries, G = 15913 rows split N ways): timing with CUDA events, or a target for ncu (--reps 1).

    python profiles/count_sharded.py N [--reps R] [--width W]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ieee_b200.engine import PackedFeatures, packed_distmat, shard_bounds
from ieee_b200.metrics.rank import GalleryLabels, RankStages
from ieee_b200.testing import market1501_shaped

ap = argparse.ArgumentParser()
ap.add_argument("shards", type=int)
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--width", type=int, default=-1, help="-1: longest merged list (steady state), 0: shards * cap")
args = ap.parse_args()
N = args.shards
s = market1501_shaped(num_q=3368 * N)
dev = torch.device("cuda")
Q, G = s.qf.shape[0], s.gf.shape[0]
qp, qc = torch.from_numpy(s.q_pids).to(dev), torch.from_numpy(s.q_camids).to(dev)
q = PackedFeatures(s.qf.to(dev), "euclidean", False, "f16x3")
parts = []
for r in range(N):
    g0, g1 = shard_bounds(G, N, r)
    g = PackedFeatures(s.gf[g0:g1].to(dev), "euclidean", False, "f16x3")
    d = torch.empty((Q, (g1 - g0 + 31) // 32 * 32), device=dev)[:, : g1 - g0]
    packed_distmat(q, g, d)
    parts.append((g0, g1, d, GalleryLabels(s.g_pids[g0:g1].copy(), s.g_camids[g0:g1].copy(), dev)))
cap = max(p[3].list_cap(qp) for p in parts)
stages = [RankStages(Q, cap, N, dev) for _ in parts]
for st, (g0, g1, d, gal) in zip(stages, parts):
    st.gather(d, qp, qc, gal, g0)
rel_all = torch.stack([st.rel for st in stages]).contiguous()
g0, g1, d, gal = parts[0]
stages[0].count(d, g1 - g0, g0, rel_all)
longest = int(stages[0].flags[2].item())
width = longest if args.width < 0 else args.width
st = RankStages(Q, cap, N, dev, width)
st.gather(d, qp, qc, gal, g0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ts = []
for _ in range(args.reps):
    flush.zero_()
    torch.cuda.synchronize()
    a.record()
    st.count(d, g1 - g0, g0, rel_all)
    b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
ms = min(ts)
print("shards %d  Q %d  local G %d  cap %d  longest merged list %d  row width %d: rank_count %.4f ms  %.0f GB/s"
      % (N, Q, g1 - g0, cap, longest, st.width, ms, 4.0 * Q * (g1 - g0) / ms / 1e6))
