"""Config 4 (BASELINE.json): synthetic 100k query x 1M gallery 3-modal embeddings, gallery sharded over N GPUs.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        profiles/c4_sharded.py [--queries 100000] [--gallery 1000000] [--check 256]

Features are generated on the device (identity-clustered, post-ReLU, D = 2304, float32); every rank owns a
contiguous gallery slice; queries are replicated.  Prints one JSON line with end-to-end queries/s (features resident
in HBM -> (cmc, mAP) on the host) and, with --check, verifies size-independent properties: the per-query AP and
first-hit rank of the first `check` queries from the SHARDED run are bit-identical to a single-GPU evaluation of the
same queries against the all-gathered gallery."""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ieee_b200.engine import RetrievalEvaluator, shard_bounds

ap = argparse.ArgumentParser()
ap.add_argument("--queries", type=int, default=100000)
ap.add_argument("--gallery", type=int, default=1000000)
ap.add_argument("--pids", type=int, default=100000)
ap.add_argument("--cams", type=int, default=8)
ap.add_argument("--dim", type=int, default=2304)
ap.add_argument("--sigma", type=float, default=2.2)
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--check", type=int, default=256)
ap.add_argument("--precision", default="f16x3")
args = ap.parse_args()

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
group = None
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
    group = dist.group.WORLD


def features(pids, centers, gen):
    out = torch.empty(pids.numel(), args.dim, device=dev)
    for s in range(0, pids.numel(), 65536):
        e = min(pids.numel(), s + 65536)
        out[s:e] = torch.relu(centers[pids[s:e]] + args.sigma * torch.randn(e - s, args.dim, device=dev, generator=gen))
    return out


gen = torch.Generator(device=dev).manual_seed(4)                 # same on every rank: centers, queries, all labels
centers = torch.randn(args.pids, args.dim, device=dev, generator=gen)
q_pids = torch.randint(0, args.pids, (args.queries,), device=dev, generator=gen)
q_cams = torch.randint(0, args.cams, (args.queries,), device=dev, generator=gen)
g_pids_all = torch.randint(0, args.pids, (args.gallery,), device=dev, generator=gen)
g_cams_all = torch.randint(0, args.cams, (args.gallery,), device=dev, generator=gen)
qf = features(q_pids, centers, gen)
g0, g1 = shard_bounds(args.gallery, world, rank)
gen_s = torch.Generator(device=dev).manual_seed(1000 + rank)     # gallery noise differs per shard
gf = features(g_pids_all[g0:g1], centers, gen_s)
del centers
torch.cuda.synchronize()


def barrier():
    if world > 1:
        dist.barrier(group=group)
    torch.cuda.synchronize()


last = {}


def step():
    ev = last["ev"] = RetrievalEvaluator(gf, g_pids_all[g0:g1], g_cams_all[g0:g1], "euclidean", False, args.precision, 20,
                                         group=group, g_offset=g0, g_total=args.gallery)
    return ev.evaluate(qf, q_pids, q_cams)


out = step()                                                     # warm-up
barrier()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(args.steps):
    out = step()
b.record()
barrier()
ms = torch.tensor([a.elapsed_time(b) / args.steps], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX, group=group)
cmc, mAP, info = out
line = {"workload": f"C4 Q={args.queries} G={args.gallery} D={args.dim} {args.precision} euclidean", "n_gpus": world,
        "ms_per_step": float(ms.item()), "queries_per_s": args.queries / float(ms.item()) * 1e3,
        "distance_tflops_algorithmic_per_gpu": 2.0 * args.queries * (g1 - g0) * args.dim / float(ms.item()) / 1e9,
        "mAP": mAP, "rank1": float(cmc[0]), "num_valid": int(info["num_valid"]), "num_ties": int(info["num_ties"]), "cap": info["cap"],
        "exchange": last["ev"].exchange}

if args.check > 0:
    n = min(args.check, args.queries)
    if world > 1:
        sizes = [shard_bounds(args.gallery, world, r)[1] - shard_bounds(args.gallery, world, r)[0] for r in range(world)]
        parts = [torch.empty((sz, args.dim), device=dev) for sz in sizes]
        dist.all_gather(parts, gf, group=group)
        gf_all = torch.cat(parts, 0)
        del parts
    else:
        gf_all = gf
    # same centre as the sharded run (it is taken from the query set: here a different one, the first n queries)
    single = RetrievalEvaluator(gf_all, g_pids_all, g_cams_all, "euclidean", False, args.precision, 20, center=last["ev"].center)
    _, _, info1 = single.evaluate(qf[:n], q_pids[:n], q_cams[:n])
    same_first = bool(torch.equal(info1["first"], info["first"][:n]))
    same_ap = bool(torch.equal(info1["ap"], info["ap"][:n]))
    line["check"] = {"queries": n, "first_hit_identical": same_first, "ap_identical": same_ap}
    assert same_first and same_ap, "sharded result differs from the single-GPU evaluation"
if rank == 0:
    print(json.dumps(line), flush=True)
if world > 1:
    dist.barrier(group=group)
    dist.destroy_process_group()
