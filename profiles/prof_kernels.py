"""Kernel-level timing / ncu target: the Market-shaped (3368 x 15913 x 2304) contraction and rank kernels.

    python profiles/prof_kernels.py [--reps N] [--sweep]

Timing uses CUDA events on the launching stream with an L2 flush (256 MB write) between repetitions.
Under ncu, run with --reps 1.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ieee_b200 import _lib
from ieee_b200.engine import PackedFeatures, packed_distmat
from ieee_b200.metrics.rank import GalleryLabels, RankStages
from ieee_b200.testing import market1501_shaped


def timeit(fn, reps, flush):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(reps):
        flush.zero_()
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), sum(ts) / len(ts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--sweep", action="store_true")
    args = ap.parse_args()
    lib = _lib.load()
    dev = torch.device("cuda")
    s = market1501_shaped()
    qf, gf = s.qf.to(dev), s.gf.to(dev)
    Q, G, D = qf.shape[0], gf.shape[0], qf.shape[1]
    flops = 2.0 * Q * G * D
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    pitch = (G + 31) // 32 * 32
    out = torch.empty((Q, pitch), dtype=torch.float32, device=dev)[:, :G]     # 128-byte row pitch -> TMA-store epilogue
    res = {}
    packs = {p: (PackedFeatures(qf, "euclidean", False, p), PackedFeatures(gf, "euclidean", False, p)) for p in ("bf16", "f16x3")}
    configs = [(cg, dbg) for cg in (2, 1) for dbg in ((0, 1, 4) if args.sweep else (0,))]
    for cg, dbg in configs:
        lib.ieee_set_cta_group(cg)
        lib.ieee_set_debug_flags(dbg)
        for chunk in ((0, 2, 4, 9) if dbg == 0 else (0,)):
            lib.ieee_set_accum_chunk(chunk)
            for prec in ("bf16", "f16x3"):
                q, g = packs[prec]
                tmin, tavg = timeit(lambda: packed_distmat(q, g, out), args.reps, flush)
                res[f"distmat_{prec}_cg{cg}_dbg{dbg}_chunk{chunk}"] = {"ms_min": tmin, "ms_avg": tavg, "tflops_alg": flops / tmin / 1e9}
    lib.ieee_set_accum_chunk(4)
    lib.ieee_set_cta_group(2)
    lib.ieee_set_debug_flags(0)
    q, g = packs["f16x3"]
    packed_distmat(q, g, out)
    gal = GalleryLabels(s.g_pids, s.g_camids, dev)
    qp = torch.from_numpy(s.q_pids).to(dev)
    qc = torch.from_numpy(s.q_camids).to(dev)
    st = RankStages(Q, gal.list_cap(qp), 1, dev)
    st.gather(out, qp, qc, gal)
    tmin, tavg = timeit(lambda: st.count(out, G), args.reps, flush)
    res["rank_count"] = {"ms_min": tmin, "ms_avg": tavg, "gbs": 4.0 * Q * G / tmin / 1e6}
    tmin, tavg = timeit(lambda: st.gather(out, qp, qc, gal), args.reps, flush)
    res["rank_gather"] = {"ms_min": tmin, "ms_avg": tavg}
    tmin, tavg = timeit(lambda: st.finalize(G, 20), args.reps, flush)
    res["rank_finalize"] = {"ms_min": tmin, "ms_avg": tavg}
    tmin, tavg = timeit(lambda: GalleryLabels(s.g_pids, s.g_camids, dev), args.reps, flush)
    res["group_gallery(+H2D labels)"] = {"ms_min": tmin, "ms_avg": tavg}
    tmin, tavg = timeit(lambda: PackedFeatures(gf, "euclidean", False, "f16x3"), args.reps, flush)
    res["pack_gallery_f16x3"] = {"ms_min": tmin, "ms_avg": tavg, "gbs": (G * D * 8.0) / tmin / 1e6}
    for k, v in res.items():
        print(k, json.dumps({a: round(b, 4) for a, b in v.items()}))


if __name__ == "__main__":
    main()
