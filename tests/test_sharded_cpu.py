"""world_size-2 gloo test (CPU) of the gallery-sharded exchange protocol of ieee_b200/engine.py:

    gather relevant (distance, global index) pairs per shard -> all_gather -> count local kept items before each
    threshold -> all_reduce(SUM) -> positions -> CMC / mAP  ==  the single-process oracle on the whole matrix.

The per-shard stages are restated in NumPy here (the CUDA kernels need a GPU; tests/test_gpu_sharded.py checks
them against the same protocol); what this test pins is the host-side algebra: contiguous shard bounds, global
indices for the tie order, padded list exchange and the additivity of the integer counts."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ieee_b200.engine import shard_bounds
from oracle import restatement as R


def test_shard_bounds_cover_exactly():
    for n in (1, 7, 15913, 1000000):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _key(d, idx):
    """(distance, global index) as one sortable tuple array: NaN last, -0 == +0 (the CUDA key order)."""
    d = np.where(np.isnan(d), np.inf, d + 0.0)
    return np.stack([d, np.isnan(d).astype(np.float64), idx.astype(np.float64)], axis=-1)


def _worker(rank, world, port, distmat, q_pids, g_pids, q_cams, g_cams, max_rank, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    Q, G = distmat.shape
    g0, g1 = shard_bounds(G, world, rank)
    d_loc, gp, gc = distmat[:, g0:g1], g_pids[g0:g1], g_cams[g0:g1]
    # stage 1: gather
    same = gp[None, :] == q_pids[:, None]
    junk = same & (gc[None, :] == q_cams[:, None])
    rel = same & ~junk
    cap = torch.tensor([int(max(rel.sum(1).max(), junk.sum(1).max(), 1))])
    dist.all_reduce(cap, op=dist.ReduceOp.MAX)
    cap = int(cap.item())
    rel_d = np.full((Q, cap), np.inf, dtype=np.float64)
    rel_i = np.full((Q, cap), -1, dtype=np.int64)
    n_rel = rel.sum(1).astype(np.int32)
    for q in range(Q):
        idx = np.nonzero(rel[q])[0]
        rel_d[q, : idx.size], rel_i[q, : idx.size] = d_loc[q, idx], idx + g0
    # exchange 1: all_gather of the padded lists
    bufs_d = [torch.empty(Q, cap, dtype=torch.float64) for _ in range(world)]
    bufs_i = [torch.empty(Q, cap, dtype=torch.int64) for _ in range(world)]
    bufs_n = [torch.empty(Q, dtype=torch.int32) for _ in range(world)]
    dist.all_gather(bufs_d, torch.from_numpy(rel_d))
    dist.all_gather(bufs_i, torch.from_numpy(rel_i))
    dist.all_gather(bufs_n, torch.from_numpy(n_rel))
    # stage 2: count local kept items before each (sorted) threshold; last column = local junk count
    counts = torch.zeros(Q, world * cap + 1, dtype=torch.int32)
    kept = ~junk
    loc_idx = np.arange(g0, g1)
    for q in range(Q):
        td = np.concatenate([bufs_d[s][q, : bufs_n[s][q]].numpy() for s in range(world)])
        ti = np.concatenate([bufs_i[s][q, : bufs_n[s][q]].numpy() for s in range(world)])
        order = np.lexsort((ti, td))
        td, ti = td[order], ti[order]
        kd, ki = d_loc[q, kept[q]], loc_idx[kept[q]]
        for k in range(td.size):
            counts[q, k] = int(((kd < td[k]) | ((kd == td[k]) & (ki < ti[k]))).sum())
        counts[q, -1] = int(junk[q].sum())
    # exchange 2: all_reduce of the integer counts
    dist.all_reduce(counts)
    n_tot = torch.stack(bufs_n).sum(0).numpy()
    positions = [np.sort(counts[q, : n_tot[q]].numpy().astype(np.int64)) for q in range(Q)]
    cmc, mAP, nv = R.metrics_from_positions(positions, max_rank, G)
    if rank == 0:
        out["cmc"], out["mAP"], out["nv"] = cmc, mAP, nv
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_protocol_equals_single_process_oracle(world):
    rng = np.random.RandomState(5)
    Q, G = 40, 301
    distmat = rng.randint(0, 60, size=(Q, G)).astype(np.float32)       # many ties, also across shard boundaries
    q_pids, g_pids = rng.randint(0, 12, Q), rng.randint(0, 12, G)
    q_cams, g_cams = rng.randint(0, 3, Q), rng.randint(0, 3, G)
    q_pids[:2] = 99                                                      # invalid queries
    cmc_o, map_o = R.evaluate_rank(distmat, q_pids, g_pids, q_cams, g_cams, max_rank=20)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), distmat, q_pids, g_pids, q_cams, g_cams, 20, out), nprocs=world, join=True)
    assert np.array_equal(out["cmc"], cmc_o) and abs(out["mAP"] - map_o) < 1e-12 and out["nv"] == Q - 2


def _topk_worker(rank, world, port, distmat, q_pids, g_pids, q_cams, g_cams, k, out):
    """Ranked lists over a sharded gallery (RetrievalEvaluator.ranked_lists): local junk-masked top-k with global
    indices -> all_gather -> merge by (distance, global index)."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    Q, G = distmat.shape
    g0, g1 = shard_bounds(G, world, rank)
    idx_l, val_l = R.topk_kept(distmat[:, g0:g1], q_pids, g_pids[g0:g1], q_cams, g_cams[g0:g1], k)
    idx_l = np.where(idx_l >= 0, idx_l + g0, -1)                          # global indices (ieee_topk's g_offset)
    bi = [torch.empty(Q, k, dtype=torch.int64) for _ in range(world)]
    bv = [torch.empty(Q, k, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(bi, torch.from_numpy(idx_l))
    dist.all_gather(bv, torch.from_numpy(val_l.astype(np.float64)))
    ci, cv = torch.cat(bi, 1).numpy(), torch.cat(bv, 1).numpy()
    idx = np.full((Q, k), -1, dtype=np.int64)
    for q in range(Q):
        live = ci[q] >= 0
        order = np.lexsort((ci[q][live], cv[q][live]))[:k]                # ieee_topk_merge: k smallest (d, index)
        idx[q, : order.size] = ci[q][live][order]
    if rank == 0:
        out["idx"] = idx
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_topk_merge_equals_single_process_oracle(world):
    rng = np.random.RandomState(9)
    Q, G, k = 30, 211, 12
    distmat = rng.randint(0, 40, size=(Q, G)).astype(np.float32)       # heavy ties across shard boundaries
    q_pids, g_pids = rng.randint(0, 10, Q), rng.randint(0, 10, G)
    q_cams, g_cams = rng.randint(0, 3, Q), rng.randint(0, 3, G)
    idx_o, _ = R.topk_kept(distmat, q_pids, g_pids, q_cams, g_cams, k)
    out = mp.Manager().dict()
    mp.spawn(_topk_worker, args=(world, _free_port(), distmat, q_pids, g_pids, q_cams, g_cams, k, out), nprocs=world, join=True)
    assert np.array_equal(out["idx"], idx_o)
