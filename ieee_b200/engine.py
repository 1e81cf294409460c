"""Device-resident evaluation: the reference's ``Engine._evaluate`` tail (torchreid/engine/engine.py:391-425)
without its host round trips.

The reference concatenates features on the CPU (engine.py:368-373), normalises (:391-394), builds the whole
Q x G matrix with torch CPU (:399-400), optionally re-ranks (:402-406) and calls ``evaluate_rank`` (:410-417).
Here features stay in HBM, the gallery is packed and grouped once, queries are processed in blocks whose
distance block lives only in HBM scratch, and -- when the process group has more than one rank -- every rank
owns a contiguous slice of the gallery: relevant-pair distances are all-gathered, integer rank counts are
all-reduced, and every rank ends with the same (cmc, mAP) the single-GPU path produces.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from . import _lib
from .metrics.rank import GalleryLabels, RankStages, _as_device, copy_stream, raise_for_status
from .utils.rerank import re_ranking_device

DEFAULT_BLOCK_BYTES = 4 << 30   # HBM scratch for one distance block


class Trace:
    """CUDA-event timeline of one evaluation (IEEE_B200_TRACE=1): mark() records an event on the current stream,
    report() -- after a synchronize -- lists the GPU time of every mark relative to the first and the host time at
    which it was enqueued.  The reference only has wall-clock AverageMeters (engine.py:365-367)."""

    def __init__(self):
        import os
        self.enabled = os.environ.get("IEEE_B200_TRACE", "0") == "1"
        self.marks = []

    def mark(self, name: str):
        if self.enabled:
            import time
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self.marks.append((name, ev, time.perf_counter()))

    def report(self):
        if not self.enabled or not self.marks:
            return
        torch.cuda.synchronize()
        _, e0, t0 = self.marks[0]
        for name, ev, t in self.marks:
            print("[trace] %-28s gpu %8.3f ms   enqueued at %8.3f ms" % (name, e0.elapsed_time(ev), (t - t0) * 1e3))
        self.marks = []


TRACE = Trace()

# List-capacity memo.  The per-query list capacity (largest identity group among the queried ids) sizes buffers and
# shared memory, and finding it costs a host round trip (plus, with several ranks, a collective that has to line the
# ranks up in the middle of a step).  It only depends on the label tensors, so it is remembered per
# (gallery ids, query ids) tensor identity AND version counter: handing in the same, unmodified tensors again skips the
# query.  It is a hint, not a trusted value: the gather kernel flags any list that does not fit, and evaluate()
# then recomputes the capacity and runs again.
_CAP_MEMO = {}
# The key is built from tensor identities (address, size, version counter), so every entry also KEEPS its key
# tensors alive: a freed label tensor whose address is reused by another tensor of the same size could otherwise
# produce a false hit -- harmless on one GPU (the overflow flag catches it), but with several ranks a hit on one
# rank and a miss on another would make them issue different collectives.
_CAP_MEMO_REFS = {}


def _memo_put(key, value, refs):
    if len(_CAP_MEMO) > 64:
        _CAP_MEMO.clear()
        _CAP_MEMO_REFS.clear()
    _CAP_MEMO[key] = value
    _CAP_MEMO_REFS[key] = refs


def _memo_drop(key):
    _CAP_MEMO.pop(key, None)
    _CAP_MEMO_REFS.pop(key, None)


_STAGE_POOL = {}


def _stages_for(Qb, cap, world, device, width=0):
    """Stage buffers are reused across evaluations of the same shape (all work on them is ordered by the compute
    stream); creating the dozen small tensors costs more host time than the kernels take to launch."""
    key = (Qb, cap, world, str(device), width)
    st = _STAGE_POOL.get(key)
    if st is None:
        if len(_STAGE_POOL) > 8:
            _STAGE_POOL.clear()
        st = _STAGE_POOL[key] = RankStages(Qb, cap, world, device, width)
    return st


_FUSED_POOL = {}
_FUSED_SKIP = {}          # label identities for which the fused path could not certify its result (values: references)
_RESULT_HOST = {}


def _result_buffer(device, nbytes):
    """Pinned host landing zone for the (cmc, summary, stats) read-back of one evaluation."""
    key = (str(device), nbytes)
    buf = _RESULT_HOST.get(key)
    if buf is None:
        buf = _RESULT_HOST[key] = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    return buf


def _fused_buffers(Q, D, precision, cap, k_eff, device):
    """Workspace + result block of the one-call path, reused per shape (stream-ordered like the stage pool)."""
    key = (Q, D, precision, cap, k_eff, str(device))
    buf = _FUSED_POOL.get(key)
    if buf is None:
        if len(_FUSED_POOL) > 8:
            _FUSED_POOL.clear()
        lib = _lib.load()
        ws = torch.empty(lib.ieee_retrieve_prepared_workspace_bytes(Q, D, precision, cap), dtype=torch.uint8, device=device)
        res_off = (4 * k_eff + 7) // 8 * 8                      # [cmc float32[k_eff] | pad | ieee_eval_summary]
        res = torch.empty(res_off + C.sizeof(_lib.EvalSummary), dtype=torch.uint8, device=device)
        res_host = torch.empty(res.shape, dtype=torch.uint8, pin_memory=True)
        buf = _FUSED_POOL[key] = (ws, res, res_off, res_host)
    return buf


def _features_key(t: torch.Tensor):
    return (t.data_ptr(), t._version, tuple(t.shape), t.stride(0), t.dtype)


def _tensor_key(t):
    return (t.data_ptr(), t.numel(), t._version, str(t.device)) if isinstance(t, torch.Tensor) else None


def shard_bounds(num_rows: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous gallery slice of ``rank``: [start, stop).  Global index = local index + start, so the
    (distance, index) tie order is the same as on one GPU."""
    base, rem = divmod(num_rows, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def centering_applies(metric: str, precision: str) -> bool:
    """Euclidean operands are packed relative to a common centre (ieee_pack_features); cosine is not translation
    invariant and the 1-pass bf16 mode multiplies bf16 inputs exactly as they are."""
    return metric == "euclidean" and precision != "bf16" and _lib.load().ieee_set_centering(-1) != 0


def feature_center(feats: torch.Tensor, normalize: bool = False, max_rows: int = 0) -> torch.Tensor:
    """Column mean over a strided sample of the rows (unit length when the rows will be normalised): float32 [D] on
    the device, deterministic (ieee_feature_center)."""
    assert feats.is_cuda and feats.dim() == 2 and feats.dtype in _lib.DTYPES and feats.shape[0] > 0
    if feats.stride(1) != 1:
        feats = feats.contiguous()
    rows, D = feats.shape
    lib = _lib.load()
    center = torch.empty(D, dtype=torch.float32, device=feats.device)
    ws = torch.empty(lib.ieee_feature_center_workspace_bytes(D), dtype=torch.uint8, device=feats.device)
    with torch.cuda.device(feats.device):
        _lib.call("ieee_feature_center", feats.data_ptr(), _lib.DTYPES[feats.dtype], feats.stride(0), rows, D, int(normalize),
                  max_rows, center.data_ptr(), ws.data_ptr(), _lib.stream())
    return center


class PackedFeatures:
    """Feature rows in the operand layout of the tensor-core kernel (ieee_pack_features)."""

    def __init__(self, feats: torch.Tensor, metric: str, normalize: bool, precision: str, center: torch.Tensor | None = None):
        if not (feats.is_cuda and feats.dim() == 2):
            raise TypeError("PackedFeatures: expected a 2-D CUDA tensor")
        if feats.dtype not in _lib.DTYPES:
            raise TypeError("PackedFeatures: features must be float32 or bfloat16, got {}".format(feats.dtype))
        if feats.stride(1) != 1:
            feats = feats.contiguous()
        self.rows, self.D = feats.shape
        self.metric, self.precision = _lib.METRICS[metric], _lib.PRECISIONS[precision]
        self.center = center
        lib = _lib.load()
        self.buf = torch.empty(max(lib.ieee_packed_bytes(self.rows, self.D, self.precision), 256), dtype=torch.uint8,
                               device=feats.device)
        if self.rows:
            _lib.call("ieee_pack_features", feats.data_ptr(), _lib.DTYPES[feats.dtype], feats.stride(0), self.rows, self.D,
                      self.metric, int(normalize), self.precision, _lib.ptr(center), self.buf.data_ptr(), _lib.stream())


def packed_distmat(q: PackedFeatures, g: PackedFeatures, out: torch.Tensor) -> torch.Tensor:
    assert q.D == g.D and q.metric == g.metric and q.precision == g.precision
    assert _lib.ptr(q.center) == _lib.ptr(g.center), "both operands must be packed with the same centre"
    _lib.call("ieee_distmat_packed", q.buf.data_ptr(), q.rows, g.buf.data_ptr(), g.rows, q.D, q.metric, q.precision,
              out.data_ptr(), out.stride(0), _fixup_workspace(q.rows, out.device).data_ptr(), _lib.stream())
    return out


_FIXUP_WS = {}


def _fixup_workspace(rows: int, device) -> torch.Tensor:
    """Near-duplicate list of the contraction (ieee_distmat_fixup_bytes): one per device and stream, grown on demand;
    contractions on one stream run one after the other, so they can share it."""
    key = (str(device), torch.cuda.current_stream(device).cuda_stream)
    need = _lib.load().ieee_distmat_fixup_bytes(rows)
    ws = _FIXUP_WS.get(key)
    if ws is None or ws.numel() < need:
        ws = _FIXUP_WS[key] = torch.empty(max(need, 1 << 16), dtype=torch.uint8, device=device)
    return ws


def _as_features(t: torch.Tensor) -> torch.Tensor:
    """float32 / bfloat16 features pass through; float16 (AMP extraction) and float64 are computed in float32, as
    compute_distance_matrix does -- there is no float64 contraction and no gradient on this path."""
    if not isinstance(t, torch.Tensor) or t.dim() != 2:
        raise TypeError("features must be a 2-D torch.Tensor")
    if t.dtype in _lib.DTYPES:
        return t.detach()
    if t.dtype in (torch.float16, torch.float64):
        return t.detach().float()
    raise TypeError("features must be a floating point tensor, got {}".format(t.dtype))


class RetrievalEvaluator:
    """distmat -> rank -> CMC/mAP over a (possibly sharded) gallery.

    ``gf, g_pids, g_camids`` are THIS rank's gallery slice (the whole gallery when ``group`` is None);
    ``g_offset`` its first global row and ``g_total`` the gallery size over all ranks.
    """

    def __init__(self, gf: torch.Tensor, g_pids, g_camids, dist_metric: str = "euclidean", normalize_feature: bool = False,
                 precision: str | None = None, max_rank: int = 20, group=None, g_offset: int = 0, g_total: int | None = None,
                 block_bytes: int = DEFAULT_BLOCK_BYTES, center: torch.Tensor | None = None, exchange: str | None = None):
        _lib.require_cuda()
        if dist_metric not in _lib.METRICS:
            raise ValueError('Unknown distance metric: {}. Please choose either "euclidean" or "cosine"'.format(dist_metric))
        self.device = gf.device if gf is not None else torch.device("cuda", torch.cuda.current_device())
        self.metric, self.normalize = dist_metric, normalize_feature
        self.precision = precision or ("bf16" if gf is not None and gf.dtype == torch.bfloat16 else "f16x3")
        self.max_rank = max_rank
        self.group = group
        self.world = 1 if group is None else torch.distributed.get_world_size(group)
        self.g_offset = g_offset
        self.block_bytes = block_bytes
        # how the ranks of a sharded gallery exchange lists and counts: "peer" = stores into NVLink peer memory from
        # inside the rank kernels (NCCL process groups), "nccl" = all-gather + all-reduce launches (any backend)
        self.exchange = exchange or os.environ.get("IEEE_B200_EXCHANGE") or (
            "peer" if self.world > 1 and torch.distributed.get_backend(group) == "nccl" else "nccl")
        if gf is not None:
            gf = _as_features(gf)
        # Euclidean operands are packed relative to a common centre (PackedFeatures).  It is taken from the QUERY set:
        # queries are replicated on every rank of a sharded gallery, so all shards -- and a single-GPU evaluation of
        # the same queries -- derive the same vector without exchanging anything.  The gallery is therefore packed
        # on the first evaluate(), or here when the caller hands a centre in.
        self._use_center = centering_applies(dist_metric, self.precision)
        self.center = center
        self._gallery_dev = gf
        # the gallery as packed row chunks [(first row, PackedFeatures)]: one chunk when the features are already
        # in HBM, several when they are streamed from the host (from_host) so that the contraction of chunk i
        # overlaps the PCIe copy of chunk i + 1
        self.chunks = []
        with torch.cuda.device(self.device):
            # a device-resident gallery is prepared by ONE foreign call (grouping on a side stream beside centre + pack):
            # here when no centre is needed or one was handed in, else in the first evaluate()
            self.labels = GalleryLabels(g_pids, g_camids, self.device, build=gf is None)
            if gf is not None and (not self._use_center or center is not None):
                self._prepare_gallery(None)
        self.G = self.labels.G
        self.g_total = self.G if g_total is None else g_total
        self._block = None
        self._early_q = None
        self._copy = None
        self._host_gallery = None
        self._label_keys = (_tensor_key(g_pids), _tensor_key(g_camids))
        self._label_refs = (g_pids, g_camids)
        self._fused_failed = False
        self.fused_stats = None

    def _prepare_gallery(self, qf_dev):
        """ieee_gallery_prepare: grouping + [centre from the query rows] + packing of a device-resident gallery -- and
        of the query rows themselves when they make one block: the host work between this call and the evaluation call
        then hides behind that kernel (a small gallery shard alone is packed before the next call arrives).  The
        grouping stays on the library's side stream until a consumer joins it (GalleryLabels.wait)."""
        gf = self._gallery_dev
        if gf.stride(1) != 1:
            gf = self._gallery_dev = gf.contiguous()
        G, D = gf.shape
        lib = _lib.load()
        prec = _lib.PRECISIONS[self.precision]
        pk = PackedFeatures.__new__(PackedFeatures)
        pk.rows, pk.D, pk.metric, pk.precision = G, D, _lib.METRICS[self.metric], prec
        pk.buf = torch.empty(max(lib.ieee_packed_bytes(G, D, prec), 256), dtype=torch.uint8, device=self.device)
        src, ws, qpk, flags = None, None, None, _lib.PREPARE_DEFER_JOIN
        if qf_dev is not None and qf_dev.shape[0] > 0:
            src = qf_dev if qf_dev.stride(1) == 1 else qf_dev.contiguous()
            assert src.dtype == gf.dtype, "query and gallery features must have the same dtype"
        if src is not None and self._use_center and self.center is None:
            self.center = torch.empty(D, dtype=torch.float32, device=self.device)
        else:
            flags |= _lib.PREPARE_KEEP_CENTER
        center = self.center if self._use_center else None
        pk.center = center
        if src is not None and src.shape[0] <= self._block_rows_for(src.shape[0], G):
            qpk = PackedFeatures.__new__(PackedFeatures)
            qpk.rows, qpk.D, qpk.metric, qpk.precision, qpk.center = src.shape[0], D, pk.metric, prec, center
            qpk.buf = torch.empty(max(lib.ieee_packed_bytes(src.shape[0], D, prec), 256), dtype=torch.uint8, device=self.device)
            self._early_q = (_features_key(qf_dev), qpk)
        cur = torch.cuda.current_stream()
        _lib.call("ieee_gallery_prepare", gf.data_ptr(), gf.stride(0), _lib.DTYPES[gf.dtype], G, D, pk.metric, int(self.normalize),
                  prec, self.labels.pids.data_ptr(), _lib.ptr(src), src.stride(0) if src is not None else 0,
                  src.shape[0] if src is not None else 0, _lib.ptr(center), pk.buf.data_ptr(),
                  self.labels.group.data_ptr(), qpk.buf.data_ptr() if qpk is not None else None, flags, _lib.ptr(ws),
                  cur.cuda_stream)
        self.labels.ready = cur.record_event()
        self.labels.side_join = True
        self.chunks = [(0, pk)]

    def _take_early_q(self, qf: torch.Tensor):
        """The packed form of exactly these query rows, if _prepare_gallery made it (used once)."""
        early, self._early_q = self._early_q, None
        if early is not None and early[0] == _features_key(qf):
            return early[1]
        return None

    def _ensure_center(self, qf_dev: torch.Tensor):
        """Fix the centre (first query set seen) and prepare a device-resident gallery with it."""
        if self._gallery_dev is not None and not self.chunks:
            self._prepare_gallery(qf_dev)
        elif self._use_center and self.center is None:
            self.center = feature_center(qf_dev, self.normalize)

    # -- one query block ---------------------------------------------------------------------------------
    def _block_rows_for(self, Q: int, G: int) -> int:
        rows = max(128, int(self.block_bytes // (4 * max(G, 1))) // 128 * 128)
        return min(Q, rows)

    def _block_rows(self, Q: int) -> int:
        return self._block_rows_for(Q, self.G)

    def _ensure_block(self, rows: int):
        if self._block is None or self._block.shape[0] < rows:
            # row pitch padded to 128 bytes: the contraction's TMA-store epilogue and the rank kernels'
            # 16-byte loads both want aligned rows (G itself is arbitrary, e.g. 15913)
            pitch = (self.G + 31) // 32 * 32
            self._block = torch.empty((rows, pitch), dtype=torch.float32, device=self.device)[:, : self.G]

    def _distance_block(self, qpk: PackedFeatures) -> torch.Tensor:
        """Distances of one packed query block against this rank's gallery rows, into the scratch block."""
        out = self._block[: qpk.rows]
        for i, (c0, gpk) in enumerate(self.chunks):
            if isinstance(gpk, tuple):            # (event, host->device staging tensor): pack on arrival
                ev, staged = gpk
                torch.cuda.current_stream().wait_event(ev)
                gpk = PackedFeatures(staged, self.metric, self.normalize, self.precision, self.center)
                self.chunks[i] = (c0, gpk)        # packed once, reused by later query blocks
            packed_distmat(qpk, gpk, out[:, c0: c0 + gpk.rows])
            TRACE.mark("  chunk %d contraction" % c0)
        return out

    def ranked_lists(self, qf: torch.Tensor, q_pids=None, q_camids=None, k: int = 10):
        """First ``k`` entries of every query's junk-filtered ranked list over the WHOLE (possibly sharded) gallery --
        what ``visualize_ranked_results`` walks (torchreid/utils/reidtools.py:49,109-145).

        Each rank selects the k best kept items of its gallery slice (``ieee_topk`` with global indices), the
        per-shard lists are all-gathered and merged (``ieee_topk_merge``): k * world candidates per query instead of
        a Q x G matrix on one device.  Ties go to the lower global gallery index, as on one GPU.  Pass no labels for an
        unmasked top-k.  Returns (idx int32 [Q, k'] global gallery indices, -1 padded; dist float32 [Q, k']) on the
        device, identical on every rank; k' = min(k, total gallery size)."""
        qf = _as_features(qf)
        masked = q_pids is not None
        with torch.cuda.device(self.device):
            qf = qf.to(self.device, non_blocking=True)
            Q = qf.shape[0]
            k_eff = max(1, min(int(k), self.g_total))
            idx = torch.empty((Q, k_eff), dtype=torch.int32, device=self.device)
            val = torch.empty((Q, k_eff), dtype=torch.float32, device=self.device)
            if Q == 0:
                return idx, val
            self._ensure_center(qf)
            self._early_q = None          # (this path packs its query blocks itself: do not keep the early copy alive)
            if self._host_gallery is not None:
                if self._copy is None:
                    self._copy = copy_stream(self.device)
                self._copy.wait_stream(torch.cuda.current_stream())
                self._start_gallery_copies(self._copy)
            qp = _as_device(q_pids, torch.int64, self.device) if masked else None
            qc = _as_device(q_camids, torch.int64, self.device) if masked else None
            rows = self._block_rows(Q)
            self._ensure_block(rows)
            self.labels.wait(torch.cuda.current_stream())
            for s in range(0, Q, rows):
                e = min(Q, s + rows)
                dist = self._distance_block(PackedFeatures(qf[s:e], self.metric, self.normalize, self.precision, self.center))
                loc_i = idx[s:e] if self.world == 1 else torch.empty((e - s, k_eff), dtype=torch.int32, device=self.device)
                loc_v = val[s:e] if self.world == 1 else torch.empty((e - s, k_eff), dtype=torch.float32, device=self.device)
                _lib.call("ieee_topk", dist.data_ptr(), dist.stride(0), e - s, self.G, self.g_offset,
                          _lib.ptr(qp[s:e] if masked else None), _lib.ptr(qc[s:e] if masked else None),
                          self.labels.pids.data_ptr() if masked else None, self.labels.camids.data_ptr() if masked else None,
                          k_eff, loc_i.data_ptr(), loc_v.data_ptr(), _lib.stream())
                if self.world > 1:
                    import torch.distributed as dist_
                    all_i = torch.empty((self.world, e - s, k_eff), dtype=torch.int32, device=self.device)
                    all_v = torch.empty((self.world, e - s, k_eff), dtype=torch.float32, device=self.device)
                    dist_.all_gather_into_tensor(all_i, loc_i, group=self.group)
                    dist_.all_gather_into_tensor(all_v, loc_v, group=self.group)
                    _lib.call("ieee_topk_merge", all_i.data_ptr(), all_v.data_ptr(), self.world, e - s, k_eff,
                              idx[s:e].data_ptr(), val[s:e].data_ptr(), _lib.stream())
        return idx, val

    def _rank_block_peer(self, dist, qp, qc, link, Qb_max, Qtot, q_base, cap, W, stats, cmc, summ, stats_out, k_eff):
        """gather -> count -> metrics of one query block with the exchanges as stores into peer memory; the metrics
        kernel of the last block also reduces over all queries."""
        Qb = dist.shape[0]
        ex = link.descriptor(Qb_max, Qb, Qtot, q_base, cap, W, link.next_epoch())
        junk = torch.empty((Qb, cap), dtype=torch.int64, device=self.device)
        n_rel = torch.empty(Qb, dtype=torch.int32, device=self.device)
        n_junk = torch.empty(Qb, dtype=torch.int32, device=self.device)
        cur = torch.cuda.current_stream()
        self.labels.wait(cur)
        st = cur.cuda_stream
        _lib.call("ieee_rank_gather_peer", dist.data_ptr(), dist.stride(0), self.G, qp.data_ptr(), qc.data_ptr(),
                  self.labels.camids.data_ptr(), self.labels.group.data_ptr(), self.g_offset, n_rel.data_ptr(), junk.data_ptr(),
                  n_junk.data_ptr(), stats.data_ptr(), C.byref(ex), st)
        TRACE.mark("  gather (lists stored into every peer)")
        _lib.call("ieee_rank_count_peer", dist.data_ptr(), dist.stride(0), self.G, self.g_offset, n_rel.data_ptr(),
                  junk.data_ptr(), n_junk.data_ptr(), stats.data_ptr(), C.byref(ex), st)
        TRACE.mark("  count (partial counts stored into every peer)")
        _lib.call("ieee_rank_metrics_peer", self.g_total, k_eff, stats.data_ptr(), cmc.data_ptr(), summ.data_ptr(),
                  stats_out.data_ptr(), C.byref(ex), st)
        TRACE.mark("  metrics (+ reduction on the last block)")
        return ex

    def _rank_block(self, dist, qp, qc, cap, width, ap, first, short, ties, inp):
        Qb = dist.shape[0]
        st = _stages_for(Qb, cap, self.world, self.device, width)
        self.labels.wait(torch.cuda.current_stream())
        st.gather(dist, qp, qc, self.labels, self.g_offset, stats=ties)     # kernels max / add straight into `ties`
        TRACE.mark("  gather")
        if self.world > 1:
            import torch.distributed as dist_
            # the lists carry their own lengths (last column): one all-gather, then one all-reduce of the counts
            rel_all = torch.empty((self.world, Qb, cap + 1), dtype=torch.int64, device=self.device)
            dist_.all_gather_into_tensor(rel_all, st.rel, group=self.group)
            TRACE.mark("  all-gather of relevant lists")
            st.count(dist, self.G, self.g_offset, rel_all)
            TRACE.mark("  count")
            dist_.all_reduce(st.counts, group=self.group)
            TRACE.mark("  all-reduce of counts")
        else:
            st.count(dist, self.G, self.g_offset)
        _lib.call("ieee_rank_query_metrics", st.counts.data_ptr(), Qb, self.g_total, 1, st.width, self.max_rank,
                  ap.data_ptr(), first.data_ptr(), short.data_ptr(), inp.data_ptr(), _lib.stream())
        TRACE.mark("  query metrics")
        return st

    @classmethod
    def from_host(cls, gf_host: torch.Tensor, g_pids, g_camids, dist_metric="euclidean", normalize_feature=False,
                  precision=None, max_rank=20, num_chunks=8, device=None, **kw):
        """Gallery features in (pinned) host memory: the copy is split into row chunks on a copy stream; each chunk is
        packed and multiplied as soon as it lands, so PCIe time and tensor-core time overlap."""
        _lib.require_cuda()
        dev = device or torch.device("cuda", torch.cuda.current_device())
        gf_host = _as_features(gf_host)
        with torch.cuda.device(dev):
            self = cls(None, g_pids, g_camids, dist_metric, normalize_feature,
                       precision or ("bf16" if gf_host.dtype == torch.bfloat16 else "f16x3"), max_rank, **kw)
        assert gf_host.shape[0] == self.G
        self._host_gallery, self._num_chunks = gf_host, num_chunks      # copies start in evaluate(), after the queries'
        return self

    def _start_gallery_copies(self, copy: torch.cuda.Stream):
        gf_host, G = self._host_gallery, self.G
        step = max(256, ((G + self._num_chunks - 1) // self._num_chunks + 255) // 256 * 256)   # whole 256-column tiles
        for c0 in range(0, G, step):
            c1 = min(G, c0 + step)
            staged = torch.empty((c1 - c0, gf_host.shape[1]), dtype=gf_host.dtype, device=self.device)
            with torch.cuda.stream(copy):
                staged.copy_(gf_host[c0:c1], non_blocking=True)
                ev = copy.record_event()
            self.chunks.append((c0, (ev, staged)))
        self._host_gallery = None

    def _evaluate_fused(self, qf, q_pids, q_camids):
        """The count fused into the contraction's epilogue (ieee_retrieve_eval_fused_prepared): no distance block, no
        capacity query.  Returns None when the result was not certified (the caller then takes the staged path, and
        this evaluator's labels are remembered as not eligible)."""
        with torch.cuda.device(self.device):
            lib = _lib.load()
            Q, D = qf.shape
            if qf.stride(1) != 1:
                qf = qf.contiguous()
            qp = _as_device(q_pids, torch.int64, self.device)
            qc = _as_device(q_camids, torch.int64, self.device)
            k_eff = min(self.max_rank, self.g_total)
            key = ("fused", Q, self.G, D, k_eff, str(self.device))
            buf = _FUSED_POOL.get(key)
            if buf is None:
                if len(_FUSED_POOL) > 8:
                    _FUSED_POOL.clear()
                ws = torch.empty(lib.ieee_retrieve_fused_workspace_bytes(Q, self.G, D), dtype=torch.uint8, device=self.device)
                res_off = (32 + 4 * k_eff + 7) // 8 * 8                 # [stats uint64[3] + pad | cmc | summary]
                res = torch.empty(res_off + C.sizeof(_lib.EvalSummary), dtype=torch.uint8, device=self.device)
                buf = _FUSED_POOL[key] = (ws, res, res_off, torch.empty(res.shape, dtype=torch.uint8, pin_memory=True),
                                          lib.ieee_retrieve_fused_spill_capacity(Q, self.G))
            ws, res, res_off, res_host, spill_cap = buf
            ap = torch.empty(Q, dtype=torch.float64, device=self.device)
            first = torch.empty(Q, dtype=torch.int32, device=self.device)
            cur = torch.cuda.current_stream()
            self.labels.wait(cur)
            gpk = self.chunks[0][1]
            _lib.call("ieee_retrieve_eval_fused_prepared", qf.data_ptr(), qf.stride(0), _lib.DTYPES[qf.dtype], Q, D,
                      _lib.METRICS[self.metric], int(self.normalize), gpk.buf.data_ptr(), self.labels.group.data_ptr(),
                      _lib.ptr(self.center), self.G, qp.data_ptr(), qc.data_ptr(), self.labels.camids.data_ptr(), self.max_rank,
                      res.data_ptr() + 32, res.data_ptr() + res_off, ap.data_ptr(), first.data_ptr(), res.data_ptr(),
                      ws.data_ptr(), ws.numel(), cur.cuda_stream)
            res_host.copy_(res, non_blocking=True)
            cur.synchronize()
            out = res_host.numpy()
            stats = out[:24].view(np.uint64)
            fallback, spilled = int(stats[0]), int(stats[2] & 0xFFFFFFFF)
            self.fused_stats = {"fallback": fallback, "spilled_spans": spilled, "spill_capacity": spill_cap,
                                "spans": Q * ((self.G + 127) // 128)}
            if fallback or spilled > spill_cap:
                return None
            cmc_host = out[32: 32 + 4 * k_eff].view(np.float32).copy()
            summary = _lib.EvalSummary.from_buffer_copy(out[res_off: res_off + 64].tobytes())
        raise_for_status(summary, self.max_rank)
        info = {"num_valid": summary.num_valid, "num_ties": summary.num_ties, "cap": 0, "ap": ap, "first": first,
                "mINP": float(summary.mINP), "fused": True}
        return cmc_host, float(summary.mAP), info

    def _evaluate_one_call(self, qf, q_pids, q_camids, return_distmat, use_cap_memo):
        """One GPU, gallery packed in HBM, queries in HBM, one distance block: the whole evaluation is ONE foreign
        call (ieee_retrieve_eval_prepared enqueues pack -> contraction -> gather -> count -> metrics -> reduce back to
        back from C) and one device->host copy of (cmc, summary)."""
        with torch.cuda.device(self.device):
            Q, D = qf.shape
            if qf.stride(1) != 1:
                qf = qf.contiguous()
            qp = _as_device(q_pids, torch.int64, self.device)
            qc = _as_device(q_camids, torch.int64, self.device)
            memo_key = None
            if use_cap_memo and self._label_keys[0] is not None and _tensor_key(q_pids) is not None:
                memo_key = (self._label_keys, _tensor_key(q_pids), self.world)
            cap = _CAP_MEMO.get(memo_key, (0, 0))[0] if memo_key is not None else 0
            cur = torch.cuda.current_stream()
            self.labels.wait(cur, join=False)
            if cap <= 0:
                # sizing pass: the capacity query synchronises once; later calls with the same labels skip it
                cap = self.labels.list_cap(qp)
                if memo_key is not None:
                    _memo_put(memo_key, (cap, 0), (self._label_refs, q_pids))
            k_eff = min(self.max_rank, self.g_total)
            prec = _lib.PRECISIONS[self.precision]
            ws, res, res_off, res_host = _fused_buffers(Q, D, prec, cap, k_eff, self.device)
            pitch = (self.G + 31) // 32 * 32
            if self._block is None or self._block.shape[0] < Q:
                self._block = torch.empty((Q, pitch), dtype=torch.float32, device=self.device)[:, : self.G]
            ap = torch.empty(Q, dtype=torch.float64, device=self.device)
            first = torch.empty(Q, dtype=torch.int32, device=self.device)
            gpk = self.chunks[0][1]
            qpk = self._take_early_q(qf)          # packed already by ieee_gallery_prepare (first evaluation of this gallery)
            _lib.call("ieee_retrieve_eval_prepared", qf.data_ptr(), qf.stride(0), _lib.DTYPES[qf.dtype], Q, D,
                      _lib.METRICS[self.metric], int(self.normalize), prec, gpk.buf.data_ptr(), self.labels.group.data_ptr(),
                      _lib.ptr(self.center), self.G, qp.data_ptr(), qc.data_ptr(), self.labels.camids.data_ptr(), self.max_rank, cap, None,
                      self._block.data_ptr(), self._block.stride(0), res.data_ptr(), res.data_ptr() + res_off,
                      ap.data_ptr(), first.data_ptr(), qpk.buf.data_ptr() if qpk is not None else None, ws.data_ptr(), ws.numel(),
                      cur.cuda_stream)
            res_host.copy_(res, non_blocking=True)
            cur.synchronize()
            out = res_host.numpy()
            cmc_host = out[: 4 * k_eff].view(np.float32).copy()
            summary = _lib.EvalSummary.from_buffer_copy(out[res_off: res_off + 64].tobytes())
        if summary.list_overflow:
            _memo_drop(memo_key)                 # stale hint: size the lists again (exact value, so this ends)
            return self._evaluate_one_call(qf, q_pids, q_camids, return_distmat, use_cap_memo)
        raise_for_status(summary, self.max_rank)
        info = {"num_valid": summary.num_valid, "num_ties": summary.num_ties, "cap": cap, "ap": ap, "first": first,
                "mINP": float(summary.mINP)}
        if return_distmat:
            info["distmat"] = self._block[:Q].clone()
        return cmc_host, float(summary.mAP), info

    def _evaluate_one_call_peer(self, qf, q_pids, q_camids, return_distmat, cap, width, memo_key):
        """Sharded gallery, sizes known from an earlier evaluation with the same labels: ONE foreign call enqueues the
        whole step (ieee_retrieve_eval_prepared_peer) -- no collective launch, no host round trip before the result."""
        from .peer import link_for
        with torch.cuda.device(self.device):
            lib = _lib.load()
            Q, D = qf.shape
            if qf.stride(1) != 1:
                qf = qf.contiguous()
            qp = _as_device(q_pids, torch.int64, self.device)
            qc = _as_device(q_camids, torch.int64, self.device)
            W = min(width, self.world * cap)
            k_eff = min(self.max_rank, self.g_total)
            prec = _lib.PRECISIONS[self.precision]
            link = link_for(self.group, self.device, lib.ieee_peer_exchange_bytes(Q, Q, cap, W, self.world))
            ex = link.descriptor(Q, Q, Q, 0, cap, W, link.next_epoch())
            key = ("peer", Q, D, prec, cap, k_eff, str(self.device))
            buf = _FUSED_POOL.get(key)
            if buf is None:
                if len(_FUSED_POOL) > 8:
                    _FUSED_POOL.clear()
                ws = torch.empty(lib.ieee_retrieve_prepared_peer_workspace_bytes(Q, D, prec, cap), dtype=torch.uint8, device=self.device)
                res = torch.empty(32 + 64 + 4 * k_eff, dtype=torch.uint8, device=self.device)
                buf = _FUSED_POOL[key] = (ws, res, torch.empty(res.shape, dtype=torch.uint8, pin_memory=True))
            ws, res, res_host = buf
            self._ensure_block(Q)
            # per-query results straight into fresh arrays (every rank derives all of them itself): nothing to copy out
            # of the exchange buffer after the result has arrived, when the GPU would be idle
            ap = torch.empty(Q, dtype=torch.float64, device=self.device)
            first = torch.empty(Q, dtype=torch.int32, device=self.device)
            cur = torch.cuda.current_stream()
            self.labels.wait(cur, join=False)
            gpk = self.chunks[0][1]
            qpk = self._take_early_q(qf)
            _lib.call("ieee_retrieve_eval_prepared_peer", qf.data_ptr(), qf.stride(0), _lib.DTYPES[qf.dtype], Q, D,
                      _lib.METRICS[self.metric], int(self.normalize), prec, gpk.buf.data_ptr(), self.labels.group.data_ptr(),
                      _lib.ptr(self.center), self.G, self.g_total, self.g_offset, qp.data_ptr(), qc.data_ptr(),
                      self.labels.camids.data_ptr(), self.max_rank, self._block.data_ptr(), self._block.stride(0),
                      res.data_ptr() + 96, res.data_ptr() + 32, res.data_ptr(), ap.data_ptr(), first.data_ptr(), C.byref(ex),
                      qpk.buf.data_ptr() if qpk is not None else None, ws.data_ptr(), ws.numel(), cur.cuda_stream)
            res_host.copy_(res, non_blocking=True)
            cur.synchronize()
            out = res_host.numpy()
            overflow, _, longest = (int(v) for v in out[:32].view(np.int64)[:3])
            if overflow or longest > W:
                _memo_drop(memo_key)             # sizes no longer fit these labels: size again, collectively
                return self.evaluate(qf, q_pids, q_camids, return_distmat, use_cap_memo=False, one_call=False)
            summary = _lib.EvalSummary.from_buffer_copy(out[32:96].tobytes())
            cmc_host = out[96:].view(np.float32).copy()
        raise_for_status(summary, self.max_rank)
        info = {"num_valid": summary.num_valid, "num_ties": summary.num_ties, "cap": cap, "ap": ap, "first": first,
                "mINP": float(summary.mINP)}
        if return_distmat:
            info["distmat"] = self._block[:Q].clone()
        return cmc_host, float(summary.mAP), info

    def evaluate(self, qf: torch.Tensor, q_pids, q_camids, return_distmat: bool = False, use_cap_memo: bool = True,
                 one_call: bool | None = None, fused: bool | None = None):
        """Returns (cmc float32 ndarray [K'], mAP float, info dict).  Queries are replicated on every rank.
        `qf` may live in (pinned) host memory: it is copied on the copy stream ahead of the gallery chunks."""
        if one_call is None:
            one_call = os.environ.get("IEEE_B200_ONE_CALL", "1") != "0"
        qf = _as_features(qf)
        if qf.is_cuda:
            with torch.cuda.device(self.device):
                if qf.shape[0] > 0:
                    self._ensure_center(qf)      # one foreign call: grouping, centre, gallery pack -- the GPU is busy from here on
                elif self._gallery_dev is not None and not self.chunks:
                    self._use_center = False     # no query rows to take a centre from (and nothing to compare against)
                    self._prepare_gallery(None)
        if (one_call and self.world == 1 and qf.is_cuda and self._host_gallery is None and len(self.chunks) == 1
                and isinstance(self.chunks[0][1], PackedFeatures) and 0 < qf.shape[0] <= self._block_rows(qf.shape[0])):
            # opt-in (fused=True / IEEE_B200_FUSED=1): certified and bit-identical to the staged path, but on the Market-shaped
            # workload the pre-pass and the counting epilogue still cost more than the block round trip they replace
            # 
            fused = os.environ.get("IEEE_B200_FUSED", "0") == "1" if fused is None else fused
            skip_key = (self._label_keys, _tensor_key(q_pids)) if self._label_keys[0] is not None and _tensor_key(q_pids) is not None else None
            if (fused and not return_distmat and self.precision == "f16x3" and not self._fused_failed
                    and (skip_key is None or skip_key not in _FUSED_SKIP)):
                out = self._evaluate_fused(qf, q_pids, q_camids)
                if out is not None:
                    return out
                # e.g. an identity with more than 32 gallery items: these labels take the staged path from now on
                self._fused_failed = True
                if skip_key is not None:
                    if len(_FUSED_SKIP) > 64:
                        _FUSED_SKIP.clear()
                    _FUSED_SKIP[skip_key] = (self._label_refs, q_pids)
            return self._evaluate_one_call(qf, q_pids, q_camids, return_distmat, use_cap_memo)
        if (one_call and use_cap_memo and self.world > 1 and self.exchange == "peer" and qf.is_cuda and self._host_gallery is None
                and len(self.chunks) == 1 and isinstance(self.chunks[0][1], PackedFeatures)
                and 0 < qf.shape[0] <= self._block_rows(qf.shape[0]) and self._label_keys[0] is not None
                and _tensor_key(q_pids) is not None):
            memo_key = (self._label_keys, _tensor_key(q_pids), self.world)
            cap, width = _CAP_MEMO.get(memo_key, (None, 0))
            if cap is not None and width > 0:
                return self._evaluate_one_call_peer(qf, q_pids, q_camids, return_distmat, cap, width, memo_key)
        with torch.cuda.device(self.device):
            TRACE.mark("evaluate: start")
            Q = qf.shape[0]
            rows = self._block_rows(Q)
            # queries already in HBM: pack the first block before anything else, so the GPU is busy while the host
            # prepares the rest of the step
            early_pack = None
            if qf.is_cuda:
                early_pack = self._take_early_q(qf) if rows >= Q else None
                if early_pack is None:
                    early_pack = PackedFeatures(qf[: min(Q, rows)], self.metric, self.normalize, self.precision, self.center)
            # small label copies go FIRST: host->device transfers of every stream share one copy engine queue, so a
            # label copy issued after the feature copies would hold the compute stream until they have all landed
            qp = _as_device(q_pids, torch.int64, self.device)
            qc = _as_device(q_camids, torch.int64, self.device)
            ids_ready = torch.cuda.current_stream().record_event()
            q_event = None
            pending_gallery = getattr(self, "_host_gallery", None) is not None
            if not qf.is_cuda or pending_gallery:
                if self._copy is None:
                    self._copy = copy_stream(self.device)
                self._copy.wait_stream(torch.cuda.current_stream())
                if not qf.is_cuda:                       # queries first: every contraction needs them
                    if self.world > 1:
                        # the query set is the same on every rank: each copies 1/N of it over its own PCIe link and
                        # the slices are all-gathered over NVLink instead of N full host->device copies
                        import torch.distributed as dist_
                        rank = dist_.get_rank(self.group)
                        per = (Q + self.world - 1) // self.world
                        q_all = torch.empty((self.world * per, qf.shape[1]), dtype=qf.dtype, device=self.device)
                        mine = q_all[rank * per: (rank + 1) * per]
                        lo, hi = min(Q, rank * per), min(Q, (rank + 1) * per)
                        with torch.cuda.stream(self._copy):
                            if hi > lo:
                                mine[: hi - lo].copy_(qf[lo:hi], non_blocking=True)
                            if hi - lo < per:
                                mine[hi - lo:].zero_()
                            dist_.all_gather_into_tensor(q_all, mine, group=self.group)
                            q_event = self._copy.record_event()
                        qf = q_all[:Q]
                    else:
                        q_dev = torch.empty(qf.shape, dtype=qf.dtype, device=self.device)
                        with torch.cuda.stream(self._copy):
                            q_dev.copy_(qf, non_blocking=True)
                            q_event = self._copy.record_event()
                        qf = q_dev
                if pending_gallery:
                    self._start_gallery_copies(self._copy)
                if q_event is not None and Q > 0:
                    # host queries: the centre comes from their device copy, before any chunk is packed
                    torch.cuda.current_stream().wait_event(q_event)
                    self._ensure_center(qf)
            ap = torch.empty(Q, dtype=torch.float64, device=self.device)
            first = torch.empty(Q, dtype=torch.int32, device=self.device)
            short = torch.empty(Q, dtype=torch.int32, device=self.device)
            inp = torch.empty(Q, dtype=torch.float64, device=self.device)
            # one result block = one device->host copy: [stats int64[4] | summary (64 B) | cmc float32[K']];
            # stats = [gather overflow = needed capacity (int32, 0: all lists fitted), tie pairs, longest merged list, pad]
            k_eff = min(self.max_rank, self.g_total)
            res = torch.zeros(32 + 64 + 4 * k_eff, dtype=torch.uint8, device=self.device)
            ties = res[:32].view(torch.int64)
            self._ensure_block(rows)

            def contraction(s, e):
                return self._distance_block(qf_packed(s, e))

            def qf_packed(s, e):
                if s == 0 and early_pack is not None:
                    return early_pack
                if q_event is not None:
                    torch.cuda.current_stream().wait_event(q_event)
                return PackedFeatures(qf[s:e], self.metric, self.normalize, self.precision, self.center)

            memo_key = None
            if use_cap_memo and self._label_keys[0] is not None and _tensor_key(q_pids) is not None:
                memo_key = (self._label_keys, _tensor_key(q_pids), self.world)
            cap, width = _CAP_MEMO.get(memo_key, (None, 0)) if memo_key is not None else (None, 0)
            if cap is not None:
                dist = contraction(0, min(Q, rows))
                TRACE.mark("contraction(block 0) queued (capacity memo hit)")
            else:
                # The contraction of the first block is queued FIRST; the capacity is then queried on a side stream
                # that only waits for the gallery grouping and the query ids, so its host round trip (and the
                # host-side allocations below) hide behind the tensor-core kernel.
                dist = contraction(0, min(Q, rows))
                TRACE.mark("contraction(block 0) queued")
                cap_done, cap_host = self.labels.list_cap_async(qp, ids_ready)
                cap_done.synchronize()
                cap = max(int(cap_host.item()), 1)
                if self.world > 1:
                    # every rank must size its lists alike for the all-gather (after the contraction: a NCCL kernel
                    # beside the persistent GEMM would take an SM pair away from a tile cluster)
                    import torch.distributed as dist_
                    cap_t = torch.tensor([cap], dtype=torch.int32, device=self.device)
                    dist_.all_reduce(cap_t, op=dist_.ReduceOp.MAX, group=self.group)
                    cap = int(cap_t.item())
            full = None
            peer = self.world > 1 and self.exchange == "peer" and Q > 0
            summ, cmc = res[32:96], res[96:].view(torch.float32)
            if peer:
                from .peer import link_for
                lib = _lib.load()
                W = min(width, self.world * cap) if width > 0 else self.world * cap
                link = link_for(self.group, self.device, lib.ieee_peer_exchange_bytes(rows, Q, cap, W, self.world))
                stats = torch.zeros(4, dtype=torch.int64, device=self.device)      # this rank's own statistics
                ex = None
                for s in range(0, Q, rows):
                    e = min(Q, s + rows)
                    if s > 0:
                        dist = contraction(s, e)
                    ex = self._rank_block_peer(dist, qp[s:e], qc[s:e], link, rows, Q, s, cap, W, stats, cmc, summ, ties, k_eff)
                    if return_distmat:
                        full = dist.clone() if full is None else torch.cat([full, dist], 0)
                TRACE.mark("rank stages done")
                offs = [lib.ieee_peer_result_offset(i, rows, Q, cap, W, self.world) for i in (0, 1)]
                ap = link.view[offs[0]: offs[0] + 8 * Q].view(torch.float64)       # every rank holds all per-query results
                first = link.view[offs[1]: offs[1] + 4 * Q].view(torch.int32)
            else:
                for s in range(0, Q, rows):
                    e = min(Q, s + rows)
                    if s > 0:
                        dist = contraction(s, e)
                    self._rank_block(dist, qp[s:e], qc[s:e], cap, width, ap[s:e], first[s:e], short[s:e], ties, inp[s:e])
                    if return_distmat:
                        full = dist.clone() if full is None else torch.cat([full, dist], 0)
                if self.world > 1:
                    import torch.distributed as dist_
                    dist_.all_reduce(ties[:2], group=self.group)      # [2] (longest merged list) is the same on every rank
                TRACE.mark("rank stages done")
                _lib.call("ieee_rank_reduce", ap.data_ptr(), first.data_ptr(), short.data_ptr(), Q, k_eff, ties.data_ptr() + 8,
                          cmc.data_ptr(), summ.data_ptr(), inp.data_ptr(), _lib.stream())
            TRACE.mark("reduce done")
            host = _result_buffer(self.device, res.numel())                    # pinned landing zone
            host.copy_(res, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            if peer:
                ap, first = ap.clone(), first.clone()    # the exchange buffer is overwritten by the next evaluation
            out = host.numpy()
            overflow, _, longest = (int(v) for v in out[:32].view(np.int64)[:3])
            summary = _lib.EvalSummary.from_buffer_copy(out[32:96].tobytes())
            cmc_host = out[96:].view(np.float32).copy()
            overflow = overflow or (longest > (width if width > 0 else self.world * cap))   # rows narrower than a merged list
            if memo_key is not None and not overflow:
                _memo_put(memo_key, (cap, max(longest, 1)), (self._label_refs, q_pids))
        if overflow:
            # a list did not fit the (memoised) capacity: forget the hint and run again with the exact value
            if not use_cap_memo:
                raise RuntimeError("ieee_b200: per-query lists overflowed an exactly sized buffer (cap=%d, longest=%d)" % (cap, longest))
            _memo_drop(memo_key)
            return self.evaluate(qf, q_pids, q_camids, return_distmat, use_cap_memo=False, one_call=one_call)
        TRACE.report()
        raise_for_status(summary, self.max_rank)
        info = {"num_valid": summary.num_valid, "num_ties": summary.num_ties, "cap": cap, "ap": ap, "first": first,
                "mINP": float(summary.mINP)}
        if return_distmat:
            info["distmat"] = full
        return cmc_host, float(summary.mAP), info


def result_lines(cmc, mAP, ranks=(1, 5, 10, 20)):
    """The result block ``Engine._evaluate`` prints (engine.py:420-425), line by line -- the format
    tools/parse_test_res.py:70-74 parses ('mAP: 61.05%', 'Rank-1  : 89.33%', ...)."""
    lines = ["** Results **", "mAP: {:.2%}".format(mAP), "CMC curve"]
    lines += ["Rank-{:<3}: {:.2%}".format(r, cmc[r - 1]) for r in ranks if r - 1 < len(cmc)]
    return lines + ["\n"]


def evaluate(qf, gf, q_pids, g_pids, q_camids, g_camids, dist_metric="euclidean", normalize_feature=False, rerank=False,
             ranks=(1, 5, 10, 20), max_rank=20, precision=None, verbose=True, dataset_name=""):
    """Tail of ``Engine._evaluate`` (engine.py:391-441) on one GPU.  Feature tensors may live on the host
    (copied once) or on the device.  Prints the reference's result lines (engine.py:420-425), which
    tools/parse_test_res.py:70 parses, and returns (cmc, mAP)."""
    _lib.require_cuda()
    dev = qf.device if qf.is_cuda else torch.device("cuda", torch.cuda.current_device())
    streamed = (not rerank) and (not gf.is_cuda)
    if not streamed:
        qf = qf.to(dev, non_blocking=True)
        gf = gf.to(dev, non_blocking=True)
    if verbose and normalize_feature:
        print("Normalzing features with L2 norm ...")
    if verbose:
        print("Computing distance matrix with metric={} ...".format(dist_metric))
    if streamed:
        ev = RetrievalEvaluator.from_host(gf, g_pids, g_camids, dist_metric, normalize_feature, precision, max_rank)
        cmc, mAP, _ = ev.evaluate(qf, q_pids, q_camids)
    elif not rerank:
        ev = RetrievalEvaluator(gf, g_pids, g_camids, dist_metric, normalize_feature, precision, max_rank)
        cmc, mAP, _ = ev.evaluate(qf, q_pids, q_camids)
    elif rerank == "gnn":
        # alternative re-ranking mode (torchreid/utils/GPU-Re-Ranking/gnn_reranking.py:27-59; its driver L2-normalises
        # the features and uses k1 = 26, k2 = 7): the negated re-ranked similarity ranks like a distance matrix
        from .metrics.rank import evaluate_device
        from .utils.gnn_reranking import gnn_reranking_distmat
        if verbose:
            print("Applying GNN re-ranking ...")
        norm = torch.nn.functional.normalize
        distmat = gnn_reranking_distmat(norm(qf.float(), dim=1), norm(gf.float(), dim=1), 26, 7)
        cmc_t, summary, _ = evaluate_device(distmat, q_pids, g_pids, q_camids, g_camids, max_rank)
        raise_for_status(summary, max_rank)
        cmc, mAP = cmc_t.cpu().numpy(), float(summary.mAP)
    else:
        from .metrics.distance import _device_distmat
        from .metrics.rank import evaluate_device
        if verbose:
            print("Applying person re-ranking ...")
        qg = _device_distmat(qf, gf, dist_metric, normalize_feature, precision)
        qq = _device_distmat(qf, qf, dist_metric, normalize_feature, precision)
        gg = _device_distmat(gf, gf, dist_metric, normalize_feature, precision)
        distmat = re_ranking_device(qg, qq, gg)
        cmc_t, summary, _ = evaluate_device(distmat, q_pids, g_pids, q_camids, g_camids, max_rank)
        raise_for_status(summary, max_rank)
        cmc, mAP = cmc_t.cpu().numpy(), float(summary.mAP)
    if verbose:
        print("Computing CMC and mAP for {}".format(dataset_name))
        for line in result_lines(cmc, mAP, ranks):
            print(line)
    return cmc, mAP
