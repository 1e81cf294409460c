"""Generate tests/golden/visrank_small.npz from the reference's own visualize_ranked_results
(torchreid/utils/reidtools.py:18-154), imported unmodified.  BUILD container only (needs /root/reference).

    python tests/golden/make_golden_visrank.py

Synthetic 3-modality "images" (PNG, lossless) for 5 queries and 14 gallery entries, a tie-free distance matrix
(NumPy's unstable argsort then has one answer).  Stored: the inputs, and what the reference wrote -- the decoded
pixels of every grid .jpg (image mode) and the relative file list (video mode).
"""
import importlib.util
import os
import sys
import tempfile
import types

import cv2
import numpy as np

REF = "/root/reference/torchreid/utils"
HERE = os.path.dirname(os.path.abspath(__file__))


def load_reference():
    pkg = types.ModuleType("refutils")
    pkg.__path__ = [REF]
    sys.modules["refutils"] = pkg
    for name in ("tools", "reidtools"):
        spec = importlib.util.spec_from_file_location("refutils." + name, os.path.join(REF, name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules["refutils." + name] = mod
        spec.loader.exec_module(mod)
    return sys.modules["refutils.reidtools"]


def write_dataset(root, images_q, images_g, q_lab, g_lab):
    """entries: ((rgb, nir, tir) paths, pid, camid, dsetid) as the fork's datasets produce them (RGBNT201.py)."""
    def entries(prefix, images, labels):
        out = []
        for i, (img3, (pid, cam)) in enumerate(zip(images, labels)):
            paths = []
            for m, img in enumerate(img3):
                p = os.path.join(root, "%s_%03d_m%d.png" % (prefix, i, m))
                cv2.imwrite(p, img)
                paths.append(p)
            out.append((tuple(paths), int(pid), int(cam), 0))
        return out
    return entries("q", images_q, q_lab), entries("g", images_g, g_lab)


def main():
    ref = load_reference()
    rng = np.random.RandomState(7)
    Q, G, H, W, topk = 5, 14, 24, 12, 4
    images_q = rng.randint(0, 256, size=(Q, 3, H, W, 3)).astype(np.uint8)
    images_g = rng.randint(0, 256, size=(G, 3, H, W, 3)).astype(np.uint8)
    q_lab = np.array([[0, 0], [1, 1], [2, 0], [0, 1], [3, 2]])
    g_lab = np.stack([rng.randint(0, 4, G), rng.randint(0, 3, G)], 1)
    g_lab[:4] = [[0, 0], [0, 1], [1, 1], [2, 2]]            # junk and true matches for the first queries
    distmat = rng.permutation(Q * G).reshape(Q, G).astype(np.float32) / 7.0      # all distinct
    out = {"images_q": images_q, "images_g": images_g, "q_lab": q_lab, "g_lab": g_lab, "distmat": distmat,
           "topk": topk, "width": 16, "height": 32}
    with tempfile.TemporaryDirectory() as tmp:
        query, gallery = write_dataset(tmp, images_q, images_g, q_lab, g_lab)
        img_dir = os.path.join(tmp, "vis_image")
        ref.visualize_ranked_results(distmat, (query, gallery), "image", width=16, height=32, save_dir=img_dir, topk=topk)
        names = sorted(os.listdir(img_dir))
        out["image_files"] = np.array(names)
        out["image_pixels"] = np.stack([cv2.imread(os.path.join(img_dir, n)) for n in names])
        vid_dir = os.path.join(tmp, "vis_video")
        ref.visualize_ranked_results(distmat, (query, gallery), "video", save_dir=vid_dir, topk=topk)
        listing = []
        for base, _, files in os.walk(vid_dir):
            for f in files:
                listing.append(os.path.relpath(os.path.join(base, f), vid_dir))
        out["video_files"] = np.array(sorted(listing))
    np.savez_compressed(os.path.join(HERE, "visrank_small.npz"), **out)
    print("wrote visrank_small.npz:", out["image_pixels"].shape, len(out["video_files"]), "video files")


if __name__ == "__main__":
    main()
