"""Drop-in for torchreid/metrics/rank.py (Market-1501 protocol), computed on the GPU without sorting.

``evaluate_rank`` keeps the reference signature and return types (rank.py:246-287): NumPy (or torch)
arrays in, ``(cmc float32[min(max_rank, G)], mAP float)`` out.  As in this fork the Python Market-1501
protocol is what runs (rank.py:278-287, ``use_cython`` is ignored) -- here it runs in
``libieee_b200.so`` (ieee_b200/csrc/rank.cu): ties are ranked by gallery index, AP is float64.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from .. import _lib


def _as_device(x, dtype, device):
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=dtype, non_blocking=True).contiguous()
    return torch.as_tensor(np.ascontiguousarray(x)).to(device=device, dtype=dtype, non_blocking=True)


# Per-device helpers that are expensive to create (cudaStreamCreate, cudaHostAlloc) and safe to share: the side
# stream for the capacity query, the copy stream for host features, one pinned int32 for the read-back.
_SIDE_STREAMS, _COPY_STREAMS, _PINNED_CAP = {}, {}, {}


def side_stream(device) -> torch.cuda.Stream:
    key = torch.device(device).index
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=device)
    return _SIDE_STREAMS[key]


def copy_stream(device) -> torch.cuda.Stream:
    key = torch.device(device).index
    if key not in _COPY_STREAMS:
        _COPY_STREAMS[key] = torch.cuda.Stream(device=device)
    return _COPY_STREAMS[key]


class GalleryLabels:
    """Gallery ids on the device plus their pid-sorted grouping (built once, reused per query block)."""

    def __init__(self, g_pids, g_camids, device, overlap: bool = False, build: bool = True):
        """overlap=True builds the grouping on the side stream (its few tiny kernels then run beside whatever the
        caller queues next, e.g. the bandwidth-bound feature packing); consumers order themselves after `ready`.
        build=False only allocates: the caller builds the grouping itself (ieee_gallery_prepare) and sets `ready`."""
        self.pids = _as_device(g_pids, torch.int64, device)
        self.camids = _as_device(g_camids, torch.int64, device)
        self.G = self.pids.numel()
        assert self.camids.numel() == self.G
        lib = _lib.load()
        self.group = torch.empty(lib.ieee_gallery_group_bytes(self.G), dtype=torch.uint8, device=device)
        self._scratch_buf = None   # (allocated on first use: only the capacity queries need it)
        self.ready = None
        self.side_join = False     # the grouping was left on the library's side stream (ieee_gallery_prepare, deferred join)
        if not build:
            return
        with torch.cuda.device(device):
            cur = torch.cuda.current_stream()
            if overlap:
                stream = side_stream(device)
                stream.wait_stream(cur)                    # label copies / allocations queued so far
                with torch.cuda.stream(stream):
                    _lib.call("ieee_gallery_group", self.pids.data_ptr(), self.G, self.group.data_ptr(), _lib.stream())
                    self.ready = stream.record_event()     # grouping queued up to here
            else:
                _lib.call("ieee_gallery_group", self.pids.data_ptr(), self.G, self.group.data_ptr(), cur.cuda_stream)
                self.ready = cur.record_event()

    @property
    def _scratch(self) -> torch.Tensor:
        if self._scratch_buf is None:
            self._scratch_buf = torch.empty(64, dtype=torch.int32, device=self.pids.device)
        return self._scratch_buf

    def wait(self, stream: torch.cuda.Stream, join: bool = True):
        """Order `stream` after the grouping.  join=False is for the one-call C entry points, which join the library's
        side stream themselves right before their gather stage."""
        if self.ready is not None:
            stream.wait_event(self.ready)
        if join and self.side_join:
            _lib.call("ieee_gallery_group_join", stream.cuda_stream)

    def list_cap_async(self, q_pids: torch.Tensor, q_ready: torch.cuda.Event):
        """Start the capacity query on the side stream, ordered only after the grouping (`self.ready`) and the query
        ids (`q_ready`) -- NOT after whatever else the compute stream holds, so the caller can queue the contraction
        first and the host round trip hides behind it.  Returns (event, pinned int32)."""
        dev = q_pids.device
        side = side_stream(dev)
        key = dev.index
        if key not in _PINNED_CAP:
            _PINNED_CAP[key] = torch.zeros(1, dtype=torch.int32).pin_memory()
        host = _PINNED_CAP[key]
        self.wait(side)
        side.wait_event(q_ready)
        with torch.cuda.stream(side):
            _lib.call("ieee_rank_list_cap", self.group.data_ptr(), self.G, q_pids.data_ptr(), q_pids.numel(),
                      self._scratch.data_ptr(), _lib.stream())
            host.copy_(self._scratch[:1], non_blocking=True)
            done = side.record_event()
        return done, host

    def list_cap(self, q_pids: torch.Tensor) -> int:
        cap = C.c_int32(0)
        with torch.cuda.device(q_pids.device):
            self.wait(torch.cuda.current_stream())
            _lib.call("ieee_rank_list_cap_sync", self.group.data_ptr(), self.G, q_pids.data_ptr(), q_pids.numel(),
                      self._scratch.data_ptr(), C.byref(cap), _lib.stream())
        return max(int(cap.value), 1)


class RankStages:
    """The three stream-ordered stages of include/ieee_b200.h (gather / count / finalize) over torch buffers."""

    def __init__(self, Q: int, cap: int, shards: int, device, width: int = 0):
        """width: row width of the count table = longest merged (all-shard) relevant list to expect; 0 = shards * cap,
        which always fits.  flags[2] reports the longest list actually seen (see ieee_rank_count)."""
        self.Q, self.cap, self.shards, self.device = Q, cap, shards, device
        self.width = shards * cap if width <= 0 else min(width, shards * cap)
        self.rel = torch.empty((Q, cap + 1), dtype=torch.int64, device=device)  # uint64 keys; [:, cap] = list length
        self.junk = torch.empty((Q, cap), dtype=torch.int64, device=device)
        self.n_rel = torch.empty(Q, dtype=torch.int32, device=device)
        self.n_junk = torch.empty(Q, dtype=torch.int32, device=device)
        self.counts = torch.empty((Q, self.width + 2), dtype=torch.int32, device=device)
        # [0] gather overflow (int32), [1] tie pairs (uint64), [2] longest merged list (uint64)
        self.flags = torch.zeros(8, dtype=torch.int64, device=device)
        self._stats = self.flags
        self.cmc = None
        self.summary = torch.empty(C.sizeof(_lib.EvalSummary), dtype=torch.uint8, device=device)
        self.ap = torch.empty(Q, dtype=torch.float64, device=device)
        self.first = torch.empty(Q, dtype=torch.int32, device=device)
        self.ws = torch.empty(_lib.load().ieee_rank_finalize_workspace_bytes(Q), dtype=torch.uint8, device=device)

    def gather(self, distmat, q_pids, q_camids, gal: GalleryLabels, g_offset: int = 0, stats=None):
        """stats: an int64[>= 3] device tensor laid out like `flags` that the kernels max / add into (the caller zeroes
        it once and lets several query blocks accumulate); default: this object's own `flags`, zeroed here."""
        if stats is None:
            self.flags.zero_()
            stats = self.flags
        self._stats = stats
        _lib.call("ieee_rank_gather", distmat.data_ptr(), distmat.stride(0), self.Q, gal.G, q_pids.data_ptr(),
                  q_camids.data_ptr(), gal.camids.data_ptr(), gal.group.data_ptr(), g_offset, self.cap,
                  self.rel.data_ptr(), self.n_rel.data_ptr(), self.junk.data_ptr(), self.n_junk.data_ptr(),
                  stats.data_ptr(), _lib.stream())

    def count(self, distmat, G: int, g_offset: int = 0, rel_all=None):
        """rel_all: the all-gathered relevant lists [shards, Q, cap + 1] (this rank's own list on one GPU)."""
        rel_all = self.rel if rel_all is None else rel_all
        _lib.call("ieee_rank_count", distmat.data_ptr(), distmat.stride(0), self.Q, G, g_offset, self.shards, self.cap,
                  self.width, rel_all.data_ptr(), self.n_rel.data_ptr(), self.junk.data_ptr(), self.n_junk.data_ptr(),
                  self.counts.data_ptr(), self._stats.data_ptr() + 8, _lib.stream())

    def finalize(self, G_total: int, max_rank: int, counts=None, ties=None):
        """counts: the all-reduced count table (this rank's own on one GPU)."""
        counts = self.counts if counts is None else counts
        k_eff = min(max_rank, G_total)
        self.cmc = torch.empty(k_eff, dtype=torch.float32, device=self.device)
        ties_ptr = self.flags.data_ptr() + 8 if ties is None else ties.data_ptr()
        _lib.call("ieee_rank_finalize", counts.data_ptr(), self.Q, G_total, 1, self.width, max_rank, ties_ptr,
                  self.cmc.data_ptr(), self.summary.data_ptr(), self.ap.data_ptr(), self.first.data_ptr(),
                  self.ws.data_ptr(), _lib.stream())

    def read_summary(self) -> _lib.EvalSummary:
        raw = self.summary.cpu().numpy().tobytes()      # synchronises
        return _lib.EvalSummary.from_buffer_copy(raw)


def raise_for_status(s: _lib.EvalSummary, max_rank: int):
    if s.status == _lib.ERR_NO_VALID_QUERY:     # rank.py:165
        raise AssertionError("Error: all query identities do not appear in gallery")
    if s.status == _lib.ERR_SHORT_RANK_LIST:    # rank.py:150,167 would build a ragged array here
        raise ValueError("{} valid queries keep fewer than max_rank={} gallery samples after removing "
                         "same-pid/same-camid entries".format(s.num_short, max_rank))


def evaluate_device(distmat: torch.Tensor, q_pids, g_pids, q_camids, g_camids, max_rank: int,
                    gallery: GalleryLabels | None = None):
    """Device-resident evaluate_rank: returns (cmc tensor, EvalSummary, RankStages)."""
    dev = distmat.device
    with torch.cuda.device(dev):
        Q, G = distmat.shape
        qp = _as_device(q_pids, torch.int64, dev)
        qc = _as_device(q_camids, torch.int64, dev)
        gal = gallery if gallery is not None else GalleryLabels(g_pids, g_camids, dev)
        assert qp.numel() == Q and qc.numel() == Q and gal.G == G
        st = RankStages(Q, gal.list_cap(qp), 1, dev)
        st.gather(distmat, qp, qc, gal)
        st.count(distmat, G)
        st.finalize(G, max_rank)
        return st.cmc, st.read_summary(), st


# Limits of the kernels behind this module (include/ieee_b200.h); the reference has none of them, so they are checked
# here with a message instead of surfacing as an error code from the library ("Limits").
MAX_RANK_LIMIT = 8192          # rank_reduce_kernel's hit table
TOPK_LIMIT = 1024              # ieee_topk's candidate buffer


def check_limits(max_rank=None, k=None):
    if max_rank is not None and max_rank > MAX_RANK_LIMIT:
        raise ValueError("max_rank={} exceeds the {} ranks the CMC kernel tabulates".format(max_rank, MAX_RANK_LIMIT))
    if k is not None and not 1 <= k <= TOPK_LIMIT:
        raise ValueError("ranked lists hold between 1 and {} entries per query, got k={}".format(TOPK_LIMIT, k))


def _evaluate_device_f64(d, q_pids, g_pids, q_camids, g_camids, max_rank):
    dev = d.device
    with torch.cuda.device(dev):
        Q, G = d.shape
        if d.stride(1) != 1:
            d = d.contiguous()
        lab = [_as_device(x, torch.int64, dev) for x in (q_pids, g_pids, q_camids, g_camids)]
        lib = _lib.load()
        cap_bound = int(min(G, 4096))
        ws = torch.empty(lib.ieee_eval_workspace_bytes(Q, G, cap_bound), dtype=torch.uint8, device=dev)
        k_eff = min(max_rank, G)
        cmc = torch.empty(k_eff, dtype=torch.float32, device=dev)
        summ = torch.empty(C.sizeof(_lib.EvalSummary), dtype=torch.uint8, device=dev)
        _lib.call("ieee_eval_market1501_f64", d.data_ptr(), d.stride(0), Q, G, lab[0].data_ptr(), lab[1].data_ptr(),
                  lab[2].data_ptr(), lab[3].data_ptr(), max_rank, cap_bound, cmc.data_ptr(), summ.data_ptr(), ws.data_ptr(),
                  ws.numel(), _lib.stream())
        summary = _lib.EvalSummary.from_buffer_copy(summ.cpu().numpy().tobytes())
    return cmc, summary


def eval_market1501(distmat, q_pids, g_pids, q_camids, g_camids, max_rank):
    """Evaluation with market1501 metric (reference: rank.py:103-171).
    Key: for each query identity, its gallery images from the same camera view are discarded."""
    _lib.require_cuda()
    dev = distmat.device if isinstance(distmat, torch.Tensor) and distmat.is_cuda else torch.device(
        "cuda", torch.cuda.current_device())
    is_f64 = (distmat.dtype == torch.float64) if isinstance(distmat, torch.Tensor) else (np.asarray(distmat).dtype == np.float64)
    d = _as_device(distmat, torch.float64 if is_f64 else torch.float32, dev)
    assert d.dim() == 2
    num_q, num_g = d.shape
    if num_g < max_rank:
        max_rank = num_g
        print("Note: number of gallery samples is quite small, got {}".format(num_g))
    if num_q == 0 or num_g == 0:      # no query can have a kept match: the reference ends in its assertion (rank.py:165)
        raise AssertionError("Error: all query identities do not appear in gallery")
    check_limits(max_rank=max_rank)
    if is_f64:
        # rank.py:117 argsorts the matrix in the dtype it is given: float64 distances keep their float64 order
        cmc, summary = _evaluate_device_f64(d, q_pids, g_pids, q_camids, g_camids, max_rank)
        raise_for_status(summary, max_rank)
        return cmc.cpu().numpy(), float(summary.mAP)
    cmc, summary, st = evaluate_device(d, q_pids, g_pids, q_camids, g_camids, max_rank)
    if summary.status == _lib.ERR_SHORT_RANK_LIST:
        # rank.py:150,167: rows of cmc[:max_rank] stack into an array only when they are equally long -- i.e. when
        # every valid query keeps the SAME number L < max_rank of gallery items; the reference then returns L ranks
        L = _uniform_short_length(st, num_g)
        if L is not None:
            st.finalize(num_g, L)
            cmc, summary = st.cmc, st.read_summary()
    raise_for_status(summary, max_rank)
    return cmc.cpu().numpy(), float(summary.mAP)


def _uniform_short_length(st: "RankStages", num_g: int):
    """The common kept-list length of all valid queries, or None when they differ (the last two columns of the count
    rows hold the number of relevant and of junk gallery items per query)."""
    c = st.counts.cpu().numpy()
    n_rel, n_junk = c[:, -2], c[:, -1]
    kept = num_g - n_junk[n_rel > 0]
    if kept.size and int(kept.min()) == int(kept.max()) and int(kept[0]) >= 1:
        return int(kept[0])
    return None


# The fork's own evaluate_rank(use_metric_cuhk03=True) dies with a TypeError (rank.py:236-239 hands 6 arguments to the
# 8-argument eval_cuhk03): that stays the default, so a drop-in behaves like the reference.  Set this flag (or call
# eval_cuhk03 directly) to get the protocol upstream torchreid and rank_cy.pyx implement.
ENABLE_CUHK03 = False
CUHK03_MAX_GALLERY = 1024      # the junk-masked ranked list comes from ieee_topk (k <= 1024)


def eval_cuhk03(distmat, q_pids, g_pids, q_camids, g_camids, max_rank, q_timeids=None, g_timeids=None, num_repeats=10):
    """Evaluation with the cuhk03 metric (single-gallery-shot; reference: rank.py:24-100, rank_cy.pyx:37-153).
    Key: one image for each gallery identity is randomly sampled for each query identity, `num_repeats` (10) times.

    The ranking runs on the GPU: one junk-masked ranked list per query (``ieee_topk`` with k = G; the junk rule is
    same pid AND same camera, AND same time id when the fork's two extra arrays are given, rank.py:48).  The sampling
    is the reference's own: NumPy's global generator, one ``np.random.choice`` per gallery identity in order of first
    appearance in the kept ranked list (rank.py:66-72) -- seed NumPy and the result is the reference's for that seed
    (ties in the distances are ranked by gallery index; the reference leaves them to an unstable argsort)."""
    _lib.require_cuda()
    q_pids, g_pids, q_camids, g_camids = (np.asarray(a.cpu() if isinstance(a, torch.Tensor) else a).astype(np.int64)
                                          for a in (q_pids, g_pids, q_camids, g_camids))
    num_q, num_g = distmat.shape
    if num_g < max_rank:
        max_rank = num_g
        print("Note: number of gallery samples is quite small, got {}".format(num_g))
    if num_q == 0 or num_g == 0:
        raise AssertionError("Error: all query identities do not appear in gallery")
    if num_g > CUHK03_MAX_GALLERY:
        raise ValueError("eval_cuhk03: the ranked lists come from the top-k kernel, which holds at most {} gallery items per "
                         "query (got {}); the protocol is meant for CUHK03-sized galleries".format(CUHK03_MAX_GALLERY, num_g))
    qc, gc = q_camids, g_camids
    if q_timeids is not None:     # same camera AND same time id <=> same (camera, time) pair: one composite id
        qt, gt = np.asarray(q_timeids).astype(np.int64), np.asarray(g_timeids).astype(np.int64)
        lo = min(qt.min(), gt.min())
        span = int(max(qt.max(), gt.max()) - lo) + 1
        qc, gc = q_camids * span + (qt - lo), g_camids * span + (gt - lo)
    ranked = topk_ranked_list(distmat, q_pids, g_pids, qc, gc, k=num_g)[0].cpu().numpy()
    rows, aps = [], []
    for q in range(num_q):
        kept = ranked[q][ranked[q] >= 0]
        kept_pids = g_pids[kept]
        hits = (kept_pids == q_pids[q]).astype(np.int32)
        if not hits.any():
            continue                                              # query identity does not appear in the gallery
        first_seen = {}
        for pos, pid in enumerate(kept_pids.tolist()):
            first_seen.setdefault(pid, []).append(pos)
        acc = np.zeros(max_rank, dtype=np.float32)
        for _ in range(num_repeats):
            shown = np.zeros(hits.size, dtype=bool)
            for positions in first_seen.values():                 # one gallery image per person
                shown[np.random.choice(positions)] = True
            curve = np.minimum(np.cumsum(hits[shown]), 1)[:max_rank].astype(np.float32)
            acc[: curve.size] += curve
        rows.append(acc / num_repeats)
        precision = np.cumsum(hits) / (np.arange(hits.size) + 1.0)
        aps.append(float((precision * hits).sum() / hits.sum()))
    if not rows:
        raise AssertionError("Error: all query identities do not appear in gallery")
    cmc = np.asarray(rows).astype(np.float32).sum(0) / float(len(rows))
    return cmc.astype(np.float32), float(np.mean(aps))


def evaluate_py(distmat, q_pids, g_pids, q_camids, g_camids, max_rank, use_metric_cuhk03):
    if use_metric_cuhk03:
        if ENABLE_CUHK03:
            return eval_cuhk03(distmat, q_pids, g_pids, q_camids, g_camids, max_rank)
        # rank.py:236-239 passes 6 arguments to the 8-argument eval_cuhk03, so the reference raises
        # TypeError on this branch; the drop-in does the same unless ENABLE_CUHK03 is set.
        raise TypeError("eval_cuhk03() missing 2 required positional arguments: 'g_camids' and 'max_rank' "
                        "(the reference's cuhk03 branch is unreachable; set ieee_b200.metrics.rank.ENABLE_CUHK03 = True "
                        "or call eval_cuhk03 for the single-gallery-shot protocol)")
    return eval_market1501(distmat, q_pids, g_pids, q_camids, g_camids, max_rank)


def evaluate_rank(distmat, q_pids, g_pids, q_camids, g_camids, max_rank=20, use_metric_cuhk03=False, use_cython=True):
    """Evaluates CMC rank (reference: rank.py:246-287).

    Args:
        distmat (numpy.ndarray | torch.Tensor): distance matrix of shape (num_query, num_gallery); a CUDA
            tensor is used in place (no copy).
        q_pids, g_pids, q_camids, g_camids: 1-D integer arrays (identities / camera views).
        max_rank (int, optional): maximum CMC rank to be computed. Default is 20 (rank.py:252).
        use_metric_cuhk03 (bool, optional): raises TypeError, as the reference does (rank.py:236-239).
        use_cython (bool, optional): accepted and ignored, as in the reference (rank.py:278-287).
    """
    return evaluate_py(distmat, q_pids, g_pids, q_camids, g_camids, max_rank, use_metric_cuhk03)


def topk_ranked_list(distmat, q_pids=None, g_pids=None, q_camids=None, g_camids=None, k=20, g_offset=0):
    """First k entries of every query's junk-filtered ranked list (rank.py:117 + :136-140; what
    torchreid/utils/reidtools.py:49,111 walks).  Pass no labels for an unmasked top-k.
    Returns (idx int32 [Q,k] global gallery indices, -1 padded; dist float32 [Q,k]) on the device."""
    _lib.require_cuda()
    dev = distmat.device if isinstance(distmat, torch.Tensor) and distmat.is_cuda else torch.device(
        "cuda", torch.cuda.current_device())
    check_limits(k=k)
    d = _as_device(distmat, torch.float32, dev)
    Q, G = d.shape
    idx = torch.empty((Q, k), dtype=torch.int32, device=dev)
    val = torch.empty((Q, k), dtype=torch.float32, device=dev)
    masked = q_pids is not None
    lab = [(_as_device(x, torch.int64, dev) if masked else None) for x in (q_pids, q_camids, g_pids, g_camids)]
    with torch.cuda.device(dev):
        _lib.call("ieee_topk", d.data_ptr(), d.stride(0), Q, G, g_offset, _lib.ptr(lab[0]), _lib.ptr(lab[1]),
                  _lib.ptr(lab[2]), _lib.ptr(lab[3]), k, idx.data_ptr(), val.data_ptr(), _lib.stream())
    return idx, val
