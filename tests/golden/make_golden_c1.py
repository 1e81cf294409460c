"""Generate tests/golden/c1_real_model.npz: features of the reference's OWN model (random init) for config C1.

BUILD container only (needs /root/reference; nothing is written there):

    python tests/golden/make_golden_c1.py

torchreid is imported from /root/reference with stubs for its dead / absent imports (SURVEY.md section 8c and
appendix A); ``build_model('ieee3modalPart', num_classes=171, loss='margin', pretrained=False).eval()`` is run on
synthetic 3-modality images (identity base image + noise, so that same-identity features cluster) in batches of 100
as Engine._evaluate does (engine.py:357-377).  Stored: a 300-image subset of the [B, 2304] float32 features (the
set is used as query AND gallery, as the RGBNT201 test split is), RGBNT201-like labels (30 ids, 4 cameras), and
the reference's own distance matrix / CMC / mAP for them (distance.py + rank.py, unmodified).
"""
import os
import sys
import types
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def import_reference():
    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules.setdefault(name, m)
        return sys.modules[name]
    stub("numpy.lib.function_base", _parse_input_dimensions=None, _flip_dispatcher=None, append=np.append)
    stub("numpy.lib.type_check", real=np.real)
    stub("numpy.lib.twodim_base", tri=np.tri)
    try:
        import matplotlib  # noqa: F401
        import matplotlib.pyplot  # noqa: F401
    except Exception:
        mpl = stub("matplotlib")
        mpl.pyplot = stub("matplotlib.pyplot")
    for name in ("gdown", "h5py"):
        try:
            __import__(name)
        except Exception:
            stub(name)
    sys.path.insert(0, "/root/reference")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import torchreid
    return torchreid


def main():
    torchreid = import_reference()
    from torchreid.metrics.distance import compute_distance_matrix
    from torchreid.metrics.rank import evaluate_rank
    torch.manual_seed(0)
    model = torchreid.models.build_model("ieee3modalPart", num_classes=171, loss="margin", pretrained=False).eval()
    gen = torch.Generator().manual_seed(0)
    P, N = 30, 300
    pids = np.repeat(np.arange(P), N // P)
    cams = np.tile(np.arange(4), N)[:N]
    base = torch.randn(P, 3, 3, 256, 128, generator=gen)            # [pid, modality, C, H, W]
    feats = []
    with torch.no_grad():
        for s in range(0, N, 100):                                   # batches of 100 (engine.py:357-377)
            pid_b = torch.from_numpy(pids[s:s + 100])
            imgs = base[pid_b] + 0.7 * torch.randn(len(pid_b), 3, 3, 256, 128, generator=gen)
            out = model([imgs[:, 0], imgs[:, 1], imgs[:, 2]], None)
            feats.append(out.cpu().clone())
            print("batch", s, out.shape, float(out.min()), float((out == 0).float().mean()))
    f = torch.cat(feats, 0)
    assert f.shape == (N, 2304) and f.dtype == torch.float32
    distmat = compute_distance_matrix(f, f, "euclidean").numpy()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        cmc, mAP = evaluate_rank(distmat, pids, pids, cams, cams, use_metric_cuhk03=False)
    print("reference: mAP %.4f rank-1 %.4f" % (mAP, cmc[0]))
    np.savez_compressed(os.path.join(HERE, "c1_real_model.npz"), feats=f.numpy(), pids=pids, camids=cams,
                        distmat=distmat, cmc=cmc, mAP=np.float64(mAP))


if __name__ == "__main__":
    main()
