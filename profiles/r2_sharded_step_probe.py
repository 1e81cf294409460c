"""Where a sharded step spends its time, per rank (torchrun): host enqueue time, GPU span, wait for the result.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P profiles/r2_sharded_step_probe.py
"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ieee_b200 import engine
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


def step(stamps=None):
    t0 = time.perf_counter()
    ev = RetrievalEvaluator(gf, lab[2], lab[3], "euclidean", False, None, 20, group=group, g_offset=g0, g_total=G)
    out = ev.evaluate(qf, lab[0], lab[1])
    if stamps is not None:
        stamps.append(time.perf_counter() - t0)
    return out


for _ in range(6):
    step()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
# (1) wall time per step on this rank, free running
walls = []
t_all = time.perf_counter()
for _ in range(40):
    step(walls)
t_all = time.perf_counter() - t_all
walls = np.array(walls) * 1e6
# (2) the same step with the GPU work timed by events and the host's part separated
orig_sync = torch.cuda.Stream.synchronize
host_issue, sync_wait, gpu_span = [], [], []
for _ in range(20):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    marks = {}

    def timed_sync(self, _orig=orig_sync):
        b.record(self)
        marks["issued"] = time.perf_counter()
        _orig(self)
        marks["synced"] = time.perf_counter()

    torch.cuda.Stream.synchronize = timed_sync
    t0 = time.perf_counter()
    a.record()
    step()
    torch.cuda.Stream.synchronize = orig_sync
    torch.cuda.synchronize()
    host_issue.append((marks["issued"] - t0) * 1e6)
    sync_wait.append((marks["synced"] - marks["issued"]) * 1e6)
    gpu_span.append(a.elapsed_time(b) * 1e3)
msg = ("rank %d/%d: step wall mean %.0f us (p50 %.0f, p90 %.0f, max %.0f); host issue %.0f us, wait for result %.0f us, GPU span "
       "(first kernel .. result copy) %.0f us" % (rank, world, t_all / 40 * 1e6, np.percentile(walls, 50), np.percentile(walls, 90), walls.max(),
                                                  np.mean(host_issue), np.mean(sync_wait), np.mean(gpu_span)))
if world > 1:
    msgs = [None] * world
    dist.all_gather_object(msgs, msg)
    if rank == 0:
        print("\n".join(msgs), flush=True)
    dist.barrier()
    dist.destroy_process_group()
else:
    print(msg, flush=True)
