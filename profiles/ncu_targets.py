"""One launch of every kernel of the bench step (Market-shaped), for `ncu --set full`."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ieee_b200.engine import PackedFeatures, feature_center, packed_distmat
from ieee_b200.metrics.rank import GalleryLabels, RankStages, topk_ranked_list
from ieee_b200.testing import market1501_shaped

dev = torch.device("cuda")
s = market1501_shaped()
qf, gf = s.qf.to(dev), s.gf.to(dev)
Q, G = qf.shape[0], gf.shape[0]
out = torch.empty((Q, (G + 31) // 32 * 32), device=dev)[:, :G]
lab = [torch.from_numpy(x).to(dev) for x in (s.q_pids, s.q_camids, s.g_pids, s.g_camids)]
for rep in range(2):                      # first pass warms up; profile the second (ncu -s skips the first)
    c = feature_center(qf)
    gp = PackedFeatures(gf, "euclidean", False, "f16x3", c)
    qp = PackedFeatures(qf, "euclidean", False, "f16x3", c)
    packed_distmat(qp, gp, out)
    g16, q16 = PackedFeatures(gf, "euclidean", False, "bf16"), PackedFeatures(qf, "euclidean", False, "bf16")
    packed_distmat(q16, g16, out)
    packed_distmat(qp, gp, out)
    gal = GalleryLabels(lab[2], lab[3], dev)
    st = RankStages(Q, gal.list_cap(lab[0]), 1, dev)
    st.gather(out, lab[0], lab[1], gal)
    st.count(out, G)
    st.finalize(G, 20)
    topk_ranked_list(out, lab[0], lab[2], lab[1], lab[3], k=20)
    torch.cuda.synchronize()
