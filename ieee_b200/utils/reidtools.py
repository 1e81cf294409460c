"""Ranked-list consumer: ``visualize_ranked_results`` (torchreid/utils/reidtools.py:18-154).

The reference argsorts the whole Q x G matrix on the host (reidtools.py:49) and then, per query, walks the sorted
row skipping same-pid/same-camera entries until it has ``topk`` images (:109-145).  Only those first ``topk`` kept
entries are ever used, so here they come from the junk-masked top-k kernel (``ieee_topk``: one streaming pass per
row on the GPU, ties by gallery index); the host does what is host work -- OpenCV reads / resizes / pastes for
``data_type='image'``, file copies for ``'video'`` -- and produces the files the reference produces (same names,
same pixels: tests/test_visrank.py against tests/golden/visrank_small.npz, written by the reference itself).

Same name, arguments and prints as the reference.  ``distmat`` is ranked in float32 (the engine hands over a float32
matrix, engine.py:399-400).
"""
from __future__ import absolute_import, print_function

import os
import os.path as osp
import shutil

import numpy as np

from ..metrics.rank import topk_ranked_list

__all__ = ["visualize_ranked_results", "ranked_lists"]

# layout constants of the reference's figure (reidtools.py:11-15)
GRID_SPACING, QUERY_EXTRA_SPACING, BW = 10, 90, 5
BLACK, GREEN, RED = (0, 0, 0), (0, 255, 0), (0, 0, 255)      # BGR: query frame, true match, false match


def ranked_lists(distmat, dataset, topk=10, ranked=None):
    """(idx int32 [Q, topk'], matched bool [Q, topk']) -- the gallery entries reidtools.py:109-145 would show for every
    query, in order (topk' = min(topk, G)); idx is -1 where a query keeps fewer gallery items than that.
    ranked: lists computed elsewhere -- RetrievalEvaluator.ranked_lists(...) over a gallery sharded across GPUs, where no
    rank ever holds the Q x G matrix -- as idx or (idx, dist) with GLOBAL gallery indices; distmat is then unused."""
    query, gallery = dataset
    q_pid, q_cam = (np.asarray([e[i] for e in query], dtype=np.int64) for i in (1, 2))
    g_pid, g_cam = (np.asarray([e[i] for e in gallery], dtype=np.int64) for i in (1, 2))
    k = max(1, min(int(topk), len(gallery)))
    if ranked is not None:
        idx = ranked[0] if isinstance(ranked, (tuple, list)) else ranked
        idx = (idx.cpu().numpy() if hasattr(idx, "cpu") else np.asarray(idx))[:, :k].astype(np.int32)
        assert idx.shape[0] == len(query), "ranked lists must have one row per query"
    else:
        idx = topk_ranked_list(distmat, q_pid, g_pid, q_cam, g_cam, k=k)[0].cpu().numpy()
    found = idx >= 0
    matched = np.zeros(idx.shape, dtype=bool)
    matched[found] = g_pid[idx[found]] == np.broadcast_to(q_pid[:, None], idx.shape)[found]
    return idx, matched


def _first(path):
    """Entries of the fork's datasets carry one path per modality (RGB, NIR, TIR); the figure shows the first."""
    return path[0] if isinstance(path, (tuple, list)) else path


class _Figure:
    """One query's grid: the query tile, a gap, then `topk` framed gallery tiles (reidtools.py:85-101,118-134)."""

    def __init__(self, cv2, width, height, topk):
        self.cv2, self.w, self.h = cv2, width, height
        self.canvas = np.full((height, (topk + 1) * width + topk * GRID_SPACING + QUERY_EXTRA_SPACING, 3), 255, np.uint8)

    def tile(self, path, frame):
        cv2 = self.cv2
        img = cv2.resize(cv2.imread(_first(path)), (self.w, self.h))
        img = cv2.copyMakeBorder(img, BW, BW, BW, BW, cv2.BORDER_CONSTANT, value=frame)
        return cv2.resize(img, (self.w, self.h))      # second resize: the frame ends up equally wide on every tile

    def put(self, slot, path, frame):
        x0 = 0 if slot == 0 else slot * (self.w + GRID_SPACING) + QUERY_EXTRA_SPACING
        self.canvas[:, x0:x0 + self.w, :] = self.tile(path, frame)

    def save(self, path):
        self.cv2.imwrite(path, self.canvas)


def _export(src, folder, label):
    """Video mode (reidtools.py:51-76): a tracklet (tuple of frames) becomes a sub-folder, a single image a file."""
    if isinstance(src, (tuple, list)):
        dst = osp.join(folder, label)
        os.makedirs(dst, exist_ok=True)
        for frame in src:
            shutil.copy(frame, dst)
    else:
        shutil.copy(src, osp.join(folder, label.split('_TRUE')[0].split('_FALSE')[0] + '_name_' + osp.basename(src)))


def visualize_ranked_results(distmat, dataset, data_type, width=128, height=256, save_dir='', topk=10, ranked=None):
    """Visualizes ranked results (image-reid: one grid figure per query; video-reid: one folder per query with the
    ranked tracklets).  Arguments as torchreid/utils/reidtools.py:18-39, plus ``ranked``: precomputed lists (see
    ``ranked_lists``), in which case ``distmat`` may be None."""
    num_q, num_g = (len(dataset[0]), len(dataset[1])) if distmat is None else distmat.shape
    if save_dir:
        os.makedirs(save_dir, exist_ok=True)
    print('# query: {}\n# gallery {}'.format(num_q, num_g))
    print('Visualizing top-{} ranks ...'.format(topk))
    query, gallery = dataset
    assert num_q == len(query)
    assert num_g == len(gallery)
    idx, matched = ranked_lists(distmat, dataset, topk) if ranked is None else ranked_lists(distmat, dataset, topk, ranked)
    as_image = data_type == 'image'
    if as_image:
        import cv2

    for q, entry in enumerate(query):
        stem = osp.basename(osp.splitext(_first(entry[0]))[0])
        shown = [(rank, int(g), bool(m)) for rank, (g, m) in enumerate(zip(idx[q], matched[q]), start=1) if g >= 0][:topk]
        if as_image:
            fig = _Figure(cv2, width, height, topk)
            fig.put(0, entry[0], BLACK)
            for rank, g, hit in shown:
                fig.put(rank, gallery[g][0], GREEN if hit else RED)
            fig.save(osp.join(save_dir, stem + '.jpg'))
        else:
            folder = osp.join(save_dir, stem)
            os.makedirs(folder, exist_ok=True)
            _export(entry[0], folder, 'query_top000')
            for rank, g, hit in shown:
                _export(gallery[g][0], folder, 'gallery_top' + str(rank).zfill(3) + ('_TRUE' if hit else '_FALSE'))
        if (q + 1) % 100 == 0:
            print('- done {}/{}'.format(q + 1, num_q))

    print('Done. Images have been saved to "{}" ...'.format(save_dir))
