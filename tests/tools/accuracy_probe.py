"""Numerical accuracy of the distance kernels vs float64, next to the reference's own fp32 result.

    python tests/tools/accuracy_probe.py

Prints, for each precision mode and for torch-CPU fp32 (what the reference computes, distance.py:59-64):
relative error over the pairs that are not near-duplicates, the error in units of |q|^2+|g|^2 (the
cancellation scale), and the self-distance statistics when query == gallery (RGBNT201-shaped)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from ieee_b200.metrics import compute_distance_matrix
from ieee_b200.testing import make_retrieval_set, rgbnt201_shaped
from oracle import restatement as R


def stats(got, truth, scale):
    err = np.abs(got.astype(np.float64) - truth)
    mask = truth > 1e-3 * scale
    if not mask.any():
        mask = truth > 0
    return {"max_rel(d>1e-3 scale)": float((err / np.abs(truth))[mask].max()), "mean_rel": float((err / np.abs(truth))[mask].mean()),
            "max_err/scale": float((err / scale).max()), "mean_err/scale": float((err / scale).mean()),
            "mean_signed_err/scale": float(((got.astype(np.float64) - truth) / scale).mean())}


def main():
    from ieee_b200 import _lib
    lib = _lib.load()
    sets = [("clustered 400x1500", make_retrieval_set(400, 1500, 30, 4, dim=2304, seed=7)), ("rgbnt201_shaped (q==g)", rgbnt201_shaped())]
    gpath = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "golden", "c1_real_model.npz")
    if os.path.isfile(gpath):
        f = torch.from_numpy(np.load(gpath)["feats"])

        class S:
            qf, gf = f, f
        sets.append(("c1_real_model (q==g, real IEEE3modalPart features)", S))
    for name, s in sets:
        truth = R.distance_fp64(s.qf, s.gf).numpy()
        scale = ((s.qf.double() ** 2).sum(1, keepdim=True) + (s.gf.double() ** 2).sum(1, keepdim=True).t()).numpy()
        rows = {"torch_cpu_fp32 (reference)": R.compute_distance_matrix(s.qf, s.gf).numpy()}
        for centre in (0, 1):
            lib.ieee_set_centering(centre)
            for chunk in (0, 9, 6, 4, 2, 1):
                lib.ieee_set_accum_chunk(chunk)
                rows[f"f16x3 centre={centre} accum_chunk={chunk}"] = compute_distance_matrix(
                    s.qf.cuda(), s.gf.cuda(), "euclidean", precision="f16x3").cpu().numpy()
            lib.ieee_set_accum_chunk(4)
            for prec in ("bf16", "fp32_simt"):
                rows[f"{prec} centre={centre}"] = compute_distance_matrix(s.qf.cuda(), s.gf.cuda(), "euclidean", precision=prec).cpu().numpy()
        lib.ieee_set_centering(1)
        print("==", name)
        for k, v in rows.items():
            st = stats(v, truth, scale)
            if truth.shape[0] == truth.shape[1]:
                dg = np.diag(v).astype(np.float64) / np.diag(scale)
                st["self_dist/scale min,max"] = [float(dg.min()), float(dg.max())]
            print(f"{k:36s}", json.dumps({a: (float("%.4g" % b) if not isinstance(b, list) else [float("%.4g" % x) for x in b]) for a, b in st.items()}))


if __name__ == "__main__":
    main()
