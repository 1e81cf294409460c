"""Ranked-list consumer: ``visualize_ranked_results`` (torchreid/utils/reidtools.py:18-154).

The reference argsorts the whole Q x G matrix on the host (reidtools.py:49) and then, per query, walks the sorted
row skipping same-pid/same-camera entries until it has ``topk`` images (:109-145).  Only those first ``topk`` kept
entries are ever used, so here they come from the junk-masked top-k kernel (``ieee_topk``: one streaming pass per
row on the GPU, ties by gallery index) and the host only does what is host work: reading, resizing and pasting
images with OpenCV exactly as the reference does (same constants, same border/resize order), or copying tracklet
files for ``data_type='video'``.

Same name, arguments, prints and output files as the reference.  ``distmat`` is ranked in float32 (the engine hands
over a float32 matrix, engine.py:399-400).
"""
from __future__ import absolute_import, print_function

import os
import os.path as osp
import shutil

import numpy as np

from ..metrics.rank import topk_ranked_list

__all__ = ["visualize_ranked_results", "ranked_lists"]

GRID_SPACING = 10           # reidtools.py:11-15
QUERY_EXTRA_SPACING = 90
BW = 5                      # border width
GREEN = (0, 255, 0)
RED = (0, 0, 255)


def _mkdir_if_missing(dirname):
    if dirname and not osp.exists(dirname):
        os.makedirs(dirname, exist_ok=True)


def _first_path(p):
    return p[0] if isinstance(p, (tuple, list)) else p


def ranked_lists(distmat, dataset, topk=10):
    """(idx int32 [Q, topk], matched bool [Q, topk]) -- the gallery entries reidtools.py:109-145 would show for
    every query, in order; idx is -1 where a query keeps fewer than ``topk`` gallery items."""
    query, gallery = dataset
    q_pids = np.asarray([q[1] for q in query], dtype=np.int64)
    q_cams = np.asarray([q[2] for q in query], dtype=np.int64)
    g_pids = np.asarray([g[1] for g in gallery], dtype=np.int64)
    g_cams = np.asarray([g[2] for g in gallery], dtype=np.int64)
    k = max(1, min(int(topk), len(gallery)))
    idx, _ = topk_ranked_list(distmat, q_pids, g_pids, q_cams, g_cams, k=k)
    idx = idx.cpu().numpy()
    matched = np.zeros(idx.shape, dtype=bool)
    ok = idx >= 0
    matched[ok] = g_pids[idx[ok]] == np.broadcast_to(q_pids[:, None], idx.shape)[ok]
    return idx, matched


def visualize_ranked_results(distmat, dataset, data_type, width=128, height=256, save_dir='', topk=10):
    """Visualizes ranked results (image-reid: one grid figure per query; video-reid: one folder per query with the
    ranked tracklets).  Arguments as torchreid/utils/reidtools.py:18-39."""
    import cv2

    num_q, num_g = distmat.shape
    _mkdir_if_missing(save_dir)

    print('# query: {}\n# gallery {}'.format(num_q, num_g))
    print('Visualizing top-{} ranks ...'.format(topk))

    query, gallery = dataset
    assert num_q == len(query)
    assert num_g == len(gallery)

    idx, matched_all = ranked_lists(distmat, dataset, topk)

    def _cp_img_to(src, dst, rank, prefix, matched=False):      # reidtools.py:51-76
        if isinstance(src, (tuple, list)):
            if prefix == 'gallery':
                suffix = 'TRUE' if matched else 'FALSE'
                dst = osp.join(dst, prefix + '_top' + str(rank).zfill(3)) + '_' + suffix
            else:
                dst = osp.join(dst, prefix + '_top' + str(rank).zfill(3))
            _mkdir_if_missing(dst)
            for img_path in src:
                shutil.copy(img_path, dst)
        else:
            dst = osp.join(dst, prefix + '_top' + str(rank).zfill(3) + '_name_' + osp.basename(src))
            shutil.copy(src, dst)

    def _tile(path, color):                                       # reidtools.py:85-92,118-130
        img = cv2.imread(_first_path(path))
        img = cv2.resize(img, (width, height))
        img = cv2.copyMakeBorder(img, BW, BW, BW, BW, cv2.BORDER_CONSTANT, value=color)
        return cv2.resize(img, (width, height))   # resized twice: consistent border width across images

    for q_idx in range(num_q):
        qimg_path, qpid, qcamid = query[q_idx][:3]
        qimg_path_name = _first_path(qimg_path)

        if data_type == 'image':
            grid_img = 255 * np.ones((height, (topk + 1) * width + topk * GRID_SPACING + QUERY_EXTRA_SPACING, 3),
                                     dtype=np.uint8)
            grid_img[:, :width, :] = _tile(qimg_path, (0, 0, 0))
        else:
            qdir = osp.join(save_dir, osp.basename(osp.splitext(qimg_path_name)[0]))
            _mkdir_if_missing(qdir)
            _cp_img_to(qimg_path, qdir, rank=0, prefix='query')

        for rank_idx, (g_idx, matched) in enumerate(zip(idx[q_idx], matched_all[q_idx]), start=1):
            if g_idx < 0 or rank_idx > topk:
                break
            gimg_path = gallery[g_idx][0]
            if data_type == 'image':
                start = rank_idx * width + rank_idx * GRID_SPACING + QUERY_EXTRA_SPACING
                end = (rank_idx + 1) * width + rank_idx * GRID_SPACING + QUERY_EXTRA_SPACING
                grid_img[:, start:end, :] = _tile(gimg_path, GREEN if matched else RED)
            else:
                _cp_img_to(gimg_path, qdir, rank=rank_idx, prefix='gallery', matched=bool(matched))

        if data_type == 'image':
            imname = osp.basename(osp.splitext(qimg_path_name)[0])
            cv2.imwrite(osp.join(save_dir, imname + '.jpg'), grid_img)

        if (q_idx + 1) % 100 == 0:
            print('- done {}/{}'.format(q_idx + 1, num_q))

    print('Done. Images have been saved to "{}" ...'.format(save_dir))
