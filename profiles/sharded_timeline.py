"""Timeline of one gallery-sharded evaluation step (run under torchrun, N ranks): the bench workload
(Q = 3368 * N queries replicated, G = 15913 gallery rows sharded), CUDA-event marks of rank 0.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 profiles/sharded_timeline.py
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ieee_b200 import engine
from ieee_b200.engine import RetrievalEvaluator, shard_bounds
from ieee_b200.testing import market1501_shaped

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
s = market1501_shaped(num_q=3368 * world)
g0, g1 = shard_bounds(s.gf.shape[0], world, rank)
qf, gf = s.qf.to(dev), s.gf[g0:g1].to(dev)
lab = [torch.from_numpy(x).to(dev) for x in (s.q_pids, s.q_camids, s.g_pids[g0:g1].copy(), s.g_camids[g0:g1].copy())]
report = engine.TRACE.report
engine.TRACE.report = lambda: None                 # keep the marks of two consecutive steps on one time axis
group = dist.group.WORLD if world > 1 else None
for it in range(7):
    engine.TRACE.enabled = rank == 0 and it >= 5
    if it <= 5:
        dist.barrier()
        torch.cuda.synchronize()
    engine.TRACE.mark("STEP %d begins (evaluator constructed next)" % it)
    ev = RetrievalEvaluator(gf, lab[2], lab[3], group=group, g_offset=g0, g_total=s.gf.shape[0])
    engine.TRACE.mark("evaluator constructed")
    cmc, mAP, info = ev.evaluate(qf, lab[0], lab[1])
    engine.TRACE.mark("result on the host")
report()
if rank == 0:
    print("world %d: mAP %.4f rank-1 %.4f cap %d memo %s" % (world, mAP, cmc[0], info["cap"], list(engine._CAP_MEMO.values())))
dist.destroy_process_group()
