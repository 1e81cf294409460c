"""Where one device-resident step (bench.py's step_device) spends its time: host enqueue time vs GPU time.

    python profiles/r2_step_timeline.py

Runs the Market-shaped evaluation repeatedly and prints, per step: host time until the evaluate() call returns,
GPU time between the first and the last kernel (CUDA events), and the wall time of the whole step."""
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

    def step():
        ev = RetrievalEvaluator(gf, lab[2], lab[3], "euclidean", False, None, 20)
        return ev.evaluate(qf, lab[0], lab[1])

    for _ in range(5):
        step()
    torch.cuda.synchronize()
    for i in range(6):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = lib.ieee_launch_count()
        t0 = time.perf_counter()
        a.record()
        t1 = time.perf_counter()
        ev = RetrievalEvaluator(gf, lab[2], lab[3], "euclidean", False, None, 20)
        t2 = time.perf_counter()
        out = ev.evaluate(qf, lab[0], lab[1])
        t3 = time.perf_counter()
        b.record()
        torch.cuda.synchronize()
        print("step %d: constructor %.1f us, evaluate() %.1f us (returns after the result is on the host), GPU span %.1f us, launches %d"
              % (i, (t2 - t1) * 1e6, (t3 - t2) * 1e6, a.elapsed_time(b) * 1e3, lib.ieee_launch_count() - l0))
    # host cost of the pieces, GPU idle
    t0 = time.perf_counter()
    for _ in range(200):
        ev = RetrievalEvaluator(gf, lab[2], lab[3], "euclidean", False, None, 20)
    torch.cuda.synchronize()
    print("constructor only: %.1f us each" % ((time.perf_counter() - t0) / 200 * 1e6))
    os.environ["IEEE_B200_TRACE"] = "0"


if __name__ == "__main__":
    main()
