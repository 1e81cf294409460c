"""rank_count on long rows (one C4 shard: G = 125000), for timing and ncu."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ieee_b200.engine import PackedFeatures, packed_distmat
from ieee_b200.metrics.rank import GalleryLabels, RankStages
from ieee_b200.testing import make_retrieval_set

Q, G = (int(sys.argv[1]) if len(sys.argv) > 1 else 8192), 125000
s = make_retrieval_set(Q, G, 12500, 8, dim=256, sigma=2.0, seed=3)
dev = torch.device("cuda")
q, g = PackedFeatures(s.qf.to(dev), "euclidean", False, "f16x3"), PackedFeatures(s.gf.to(dev), "euclidean", False, "f16x3")
out = torch.empty((Q, (G + 31) // 32 * 32), device=dev)[:, :G]
packed_distmat(q, g, out)
lab = [torch.from_numpy(x).to(dev) for x in (s.q_pids, s.q_camids, s.g_pids, s.g_camids)]
gal = GalleryLabels(lab[2], lab[3], dev)
st = RankStages(Q, gal.list_cap(lab[0]), 1, dev)
st.gather(out, lab[0], lab[1], gal)
st.count(out, G)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(3):
    st.count(out, G)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / 3
print("rank_count G=%d Q=%d: %.3f ms  %.0f GB/s" % (G, Q, ms, 4.0 * Q * G / ms / 1e6))
