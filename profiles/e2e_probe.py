"""End-to-end step (pinned host features in, (cmc, mAP) out) for different gallery chunk counts."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ieee_b200.engine import RetrievalEvaluator
from ieee_b200.testing import market1501_shaped

s = market1501_shaped()
dev = torch.device("cuda")
qh, gh = s.qf.pin_memory(), s.gf.pin_memory()
lab = [torch.from_numpy(x).pin_memory() for x in (s.q_pids, s.q_camids, s.g_pids, s.g_camids)]

def resident():
    q, g = qh.to(dev, non_blocking=True), gh.to(dev, non_blocking=True)
    l = [t.to(dev, non_blocking=True) for t in lab]
    return RetrievalEvaluator(g, l[2], l[3]).evaluate(q, l[0], l[1])

def streamed(k):
    return RetrievalEvaluator.from_host(gh, lab[2], lab[3], num_chunks=k).evaluate(qh, lab[0], lab[1])

def bench(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3

print("copy everything, then compute      %.3f ms" % bench(resident))
for k in (1, 2, 4, 8, 16):
    print("streamed, %2d gallery chunks         %.3f ms" % (k, bench(lambda: streamed(k))))
# CPU-side cost of one step when nothing has to be waited for: device-resident inputs
qd, gd = qh.to(dev), gh.to(dev); ld = [t.to(dev) for t in lab]
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    ev = RetrievalEvaluator(gd, ld[2], ld[3]); out = ev.evaluate(qd, ld[0], ld[1])
torch.cuda.synchronize()
print("device-resident step (wall)          %.3f ms" % ((time.perf_counter() - t0) / 10 * 1e3))

os.environ["IEEE_B200_TRACE"] = "1"
from ieee_b200 import engine
engine.TRACE.enabled = True
print("--- trace of one streamed step (4 chunks)")
torch.cuda.synchronize(); t0 = time.perf_counter()
ev = RetrievalEvaluator.from_host(gh, lab[2], lab[3], num_chunks=4)
t1 = time.perf_counter()
ev.evaluate(qh, lab[0], lab[1])
print("host: from_host %.3f ms, evaluate %.3f ms" % ((t1 - t0) * 1e3, (time.perf_counter() - t1) * 1e3))
