"""Timeline of one streamed end-to-end step: when does each gallery chunk land, when is its contraction done?"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ieee_b200.engine import PackedFeatures, packed_distmat
from ieee_b200.testing import market1501_shaped

s = market1501_shaped()
dev = torch.device("cuda")
qh, gh = s.qf.pin_memory(), s.gf.pin_memory()
G, D = gh.shape
copy = torch.cuda.Stream()
main = torch.cuda.current_stream()
out = torch.empty((qh.shape[0], (G + 31) // 32 * 32), device=dev)[:, :G]
step = 4096
for rep in range(3):
    torch.cuda.synchronize()
    t_cpu0 = time.perf_counter()
    start = torch.cuda.Event(enable_timing=True); start.record()
    copy.wait_stream(main)
    qd = torch.empty_like(qh, device=dev)
    marks = []
    with torch.cuda.stream(copy):
        qd.copy_(qh, non_blocking=True)
        eq = torch.cuda.Event(enable_timing=True); eq.record()
    chunks = []
    for c0 in range(0, G, step):
        c1 = min(G, c0 + step)
        st = torch.empty((c1 - c0, D), device=dev)
        with torch.cuda.stream(copy):
            st.copy_(gh[c0:c1], non_blocking=True)
            e = torch.cuda.Event(enable_timing=True); e.record()
        chunks.append((c0, c1, st, e))
    t_cpu1 = time.perf_counter()
    main.wait_event(eq)
    qp = PackedFeatures(qd, "euclidean", False, "f16x3")
    done = []
    for c0, c1, st, e in chunks:
        main.wait_event(e)
        gp = PackedFeatures(st, "euclidean", False, "f16x3")
        packed_distmat(qp, gp, out[:, c0:c1])
        d = torch.cuda.Event(enable_timing=True); d.record()
        done.append(d)
    t_cpu2 = time.perf_counter()
    torch.cuda.synchronize()
    if rep == 2:
        print("cpu: copies enqueued after %.3f ms, all launches enqueued after %.3f ms" % ((t_cpu1 - t_cpu0) * 1e3, (t_cpu2 - t_cpu0) * 1e3))
        print("q landed            %.3f ms" % start.elapsed_time(eq))
        for (c0, c1, st, e), d in zip(chunks, done):
            print("chunk %5d..%5d landed %.3f ms, contraction done %.3f ms" % (c0, c1, start.elapsed_time(e), start.elapsed_time(d)))
