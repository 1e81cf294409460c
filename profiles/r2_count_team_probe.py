"""rank_count_warp_kernel: warps per query (ieee_set_count_team) at the Market shape, 3368 rows x 15913 columns.

    python profiles/r2_count_team_probe.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ieee_b200 import _lib
from ieee_b200.engine import PackedFeatures, feature_center, packed_distmat
from ieee_b200.metrics.rank import GalleryLabels, RankStages
from ieee_b200.testing import market1501_shaped

dev = torch.device("cuda")
lib = _lib.load()
s = market1501_shaped(seed=1)
qf, gf = s.qf.to(dev), s.gf.to(dev)
Q, G = qf.shape[0], gf.shape[0]
out = torch.empty((Q, (G + 31) // 32 * 32), device=dev)[:, :G]
lab = [torch.from_numpy(x).to(dev) for x in (s.q_pids, s.q_camids, s.g_pids, s.g_camids)]
c = feature_center(qf)
packed_distmat(PackedFeatures(qf, "euclidean", False, "f16x3", c), PackedFeatures(gf, "euclidean", False, "f16x3", c), out)
gal = GalleryLabels(lab[2], lab[3], dev)
st = RankStages(Q, gal.list_cap(lab[0]), 1, dev)
st.gather(out, lab[0], lab[1], gal)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ref = None
for team in (1, 2, 4, 8, 1):
    lib.ieee_set_count_team(team)
    for _ in range(3):
        st.count(out, G)
    cold, warm = [], []
    for _ in range(10):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); st.count(out, G); b.record(); torch.cuda.synchronize()
        cold.append(a.elapsed_time(b) * 1e3)
    for _ in range(10):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); st.count(out, G); b.record(); torch.cuda.synchronize()
        warm.append(a.elapsed_time(b) * 1e3)
    counts = st.counts.clone()
    same = True if ref is None else bool(torch.equal(counts, ref))
    ref = counts if ref is None else ref
    print("team %d: L2 flushed min %.1f us (%.0f GB/s), back to back min %.1f us (%.0f GB/s), counts identical to team 1: %s"
          % (team, min(cold), 4.0 * Q * G / min(cold) / 1e3, min(warm), 4.0 * Q * G / min(warm) / 1e3, same), flush=True)
