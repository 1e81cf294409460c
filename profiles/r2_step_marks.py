"""In-step timeline of the bench step (Market-shaped), from CUDA events the library records after each of its launches
(ieee_set_debug_flags bit 7).  Unlike an ncu launch list the kernels run back to back with warm caches, so the
differences between marks are kernel time + launch gap as they occur inside a step.

    python profiles/r2_step_marks.py
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P profiles/r2_step_marks.py
"""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ieee_b200 import _lib
from ieee_b200.engine import RetrievalEvaluator, shard_bounds
from ieee_b200.testing import market1501_shaped

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
group = None
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
    group = dist.group.WORLD
s = market1501_shaped(seed=1, num_q=3368 * world)
G = 15913
g0, g1 = shard_bounds(G, world, rank)
qf, gf = s.qf.to(dev), s.gf[g0:g1].to(dev)
lab = [torch.from_numpy(x).to(dev) for x in (s.q_pids, s.q_camids, s.g_pids[g0:g1].copy(), s.g_camids[g0:g1].copy())]
lib = _lib.load()


def step():
    ev = RetrievalEvaluator(gf, lab[2], lab[3], "euclidean", False, None, 20, group=group, g_offset=g0, g_total=G)
    return ev.evaluate(qf, lab[0], lab[1])


for _ in range(6):
    step()
lib.ieee_set_debug_flags(128)
for _ in range(3):
    step()
rows = {}
order = []
buf = C.create_string_buffer(1 << 14)
REPS = 20
for _ in range(REPS):
    lib.ieee_debug_timeline_reset()
    step()
    n = lib.ieee_debug_timeline(buf, len(buf))
    seen = {}
    for line in buf.value.decode().strip().split("\n"):
        name, d, t = line.split("\t")
        k = seen.get(name, 0)
        seen[name] = k + 1
        key = name if k == 0 else "%s #%d" % (name, k + 1)
        if key not in rows:
            rows[key] = []
            order.append(key)
        rows[key].append((float(d), float(t)))
lib.ieee_set_debug_flags(0)
lines = ["rank %d/%d: marks of one step, median of %d steps (us after the previous mark | us since the first mark)" % (rank, world, REPS)]
for key in order:
    a = np.array(rows[key])
    lines.append("  %-50s %8.1f %9.1f" % (key, np.median(a[:, 0]), np.median(a[:, 1])))
msg = "\n".join(lines)
if world > 1:
    msgs = [None] * world
    dist.all_gather_object(msgs, msg)
    if rank == 0:
        print("\n".join(m for i, m in enumerate(msgs) if i in (0, world - 1)), flush=True)
    dist.barrier()
    dist.destroy_process_group()
else:
    print(msg, flush=True)
