"""Contraction sweep of round 2: accumulation chunk x tile raster x shape (CUDA events, L2 flushed between reps).

    python profiles/r2_gemm_sweep.py [--reps N]

Shapes: the Market-shaped block (3368 x 15913) and the block one rank of an 8-way gallery shard sees in the weak
scaling bench (26944 queries x 1989 gallery rows) -- the tall case the panel raster is for.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ieee_b200 import _lib
from ieee_b200.engine import PackedFeatures, feature_center, packed_distmat


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
    ap.add_argument("--reps", type=int, default=8)
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    lib = _lib.load()
    dev = torch.device("cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    gen = torch.Generator(device=dev).manual_seed(0)
    shapes = ((3368, 15913), (26944, 1989), (8192, 125000)) if not args.quick else ((3368, 15913), (26944, 1989))
    for Q, G in shapes:
        D = 2304
        qf = torch.relu(torch.randn(Q, D, device=dev, generator=gen) + 0.3)
        gf = torch.relu(torch.randn(G, D, device=dev, generator=gen) + 0.3)
        flops = 2.0 * Q * G * D
        pitch = (G + 31) // 32 * 32
        out = torch.empty((Q, pitch), dtype=torch.float32, device=dev)[:, :G]
        c = feature_center(qf)
        packs = {"f16x3": (PackedFeatures(qf, "euclidean", False, "f16x3", c), PackedFeatures(gf, "euclidean", False, "f16x3", c)),
                 "f16x3-uncentred": (PackedFeatures(qf, "euclidean", False, "f16x3"), PackedFeatures(gf, "euclidean", False, "f16x3")),
                 "bf16": (PackedFeatures(qf, "euclidean", False, "bf16"), PackedFeatures(gf, "euclidean", False, "bf16"))}
        for panel in ((0, 100000) if not args.quick else (0,)):
            lib.ieee_set_raster_panel(panel)
            for prec, chunks in (("f16x3", (0, 2, 3, 4, 6, 9, 12) if not args.quick else (4, 6, 9)), ("f16x3-uncentred", (4, 6)), ("bf16", (0,))):
                for chunk in chunks:
                    for dbg in ((0, 32) if prec != "bf16" else (0,)):       # 32: no near-duplicate fix-up pass
                        lib.ieee_set_accum_chunk(chunk)
                        lib.ieee_set_debug_flags(dbg)
                        q, g = packs[prec]
                        tmin, tavg = timeit(lambda: packed_distmat(q, g, out), args.reps, flush)
                        nfix = int(_fix_count(dev)) if dbg == 0 and prec != "bf16" else -1
                        print(f"Q={Q} G={G} {prec} chunk={chunk} fixup={'on' if dbg == 0 else 'off'} panel={'auto' if panel == 0 else 'one (round-1 raster)'}",
                              json.dumps({"ms_min": round(tmin, 4), "ms_avg": round(tavg, 4), "tflops_alg": round(flops / tmin / 1e9, 1),
                                          "fixup_entries": nfix}), flush=True)
        lib.ieee_set_raster_panel(0)
        lib.ieee_set_accum_chunk(6)
        lib.ieee_set_debug_flags(0)
        del qf, gf, out, packs


def _fix_count(dev):
    from ieee_b200.engine import _FIXUP_WS
    ws = next(iter(_FIXUP_WS.values()))
    return ws[:8].view(torch.int64)[0].item()


if __name__ == "__main__":
    main()
