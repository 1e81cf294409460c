"""Golden vectors for the cuhk03 (single-gallery-shot) protocol, written by the REFERENCE's compiled
rank_cy.evaluate_cy(..., use_metric_cuhk03=True) (torchreid/metrics/rank_cylib/rank_cy.pyx:37-153, built unmodified
by oracle/build_ref.py) with NumPy's global generator seeded.

    python tests/golden/make_golden_cuhk03.py      # needs /root/reference (build container only)

rank.py:24-100 (the fork's 8-argument Python version) cannot run on NumPy >= 1.24 (np.bool, :67); the Cython form is
the one that executes.  Distances are tie-free, so the reference's unstable argsort is unambiguous."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402


def main():
    assert ref.available(), "oracle/_ref is not built (python -m oracle.build_ref)"
    rng = np.random.RandomState(11)
    Q, G, P, C = 60, 400, 25, 3
    distmat = rng.permutation(Q * G).reshape(Q, G).astype(np.float32) / 7.0      # all distinct: no ties anywhere
    q_pids, g_pids = rng.randint(0, P, Q).astype(np.int64), rng.randint(0, P, G).astype(np.int64)
    q_cams, g_cams = rng.randint(0, C, Q).astype(np.int64), rng.randint(0, C, G).astype(np.int64)
    q_pids[:2] = 99                                                                # identities without gallery images
    out = {"distmat": distmat, "q_pids": q_pids, "g_pids": g_pids, "q_camids": q_cams, "g_camids": g_cams}
    for seed in (0, 1234):
        for max_rank in (5, 20):
            np.random.seed(seed)
            cmc, mAP = ref.evaluate_cy(distmat, q_pids, g_pids, q_cams, g_cams, max_rank, use_metric_cuhk03=True)
            out[f"cmc_s{seed}_k{max_rank}"] = np.asarray(cmc, dtype=np.float32)
            out[f"mAP_s{seed}_k{max_rank}"] = np.float64(mAP)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "cuhk03_small.npz"), **out)
    print({k: (v.shape if hasattr(v, "shape") and v.shape else float(v)) for k, v in out.items() if k.startswith(("cmc", "mAP"))})


if __name__ == "__main__":
    main()
