"""One fused evaluation of the Market-shaped workload (for an ncu launch list) + CUDA-event timings of variants.

    python profiles/r2_fused_probe.py [--events]
"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ieee_b200 import _lib
from ieee_b200.engine import RetrievalEvaluator
from ieee_b200.testing import market1501_shaped


def main():
    dev = torch.device("cuda")
    s = market1501_shaped(seed=1)
    qf, gf = s.qf.to(dev), s.gf.to(dev)
    lab = [torch.from_numpy(x).to(dev) for x in (s.q_pids, s.q_camids, s.g_pids, s.g_camids)]
    lib = _lib.load()

    def step(fused):
        ev = RetrievalEvaluator(gf, lab[2], lab[3], "euclidean", False, None, 20)
        out = ev.evaluate(qf, lab[0], lab[1], fused=fused)
        return ev, out

    for fused in (True, False):
        for _ in range(3):
            ev, out = step(fused)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            ev, out = step(fused)
        b.record()
        torch.cuda.synchronize()
        print("fused=%s: %.3f ms per step, mAP %.6f, stats %s" % (fused, a.elapsed_time(b) / 10, out[1], ev.fused_stats), flush=True)
    if "--events" in sys.argv:
        for chunk in (6, 9, 12, 18):
            lib.ieee_set_fused_chunk(chunk)
            for _ in range(2):
                step(True)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(10):
                ev, out = step(True)
            b.record()
            torch.cuda.synchronize()
            print("fused chunk=%d: %.3f ms per step, mAP %.6f" % (chunk, a.elapsed_time(b) / 10, out[1]), flush=True)
        lib.ieee_set_fused_chunk(0)


if __name__ == "__main__":
    main()
