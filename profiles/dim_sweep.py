"""Config 5 (BASELINE.json): bf16 feature-dim sweep, cosine, 1-pass tcgen05 contraction -- tensor roofline study.

    python profiles/dim_sweep.py [--gallery 1000000] [--queries 8192]

bf16 unit-norm-ish features, Q-block x gallery contraction only (distance kernel), CUDA events, L2 flushed.
On one GPU the gallery is 1M rows (10M x 6144 bf16 = 123 GB does not leave room for the distance blocks);
time per (query, gallery) pair is independent of G at this size, so TFLOP/s carries over."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ieee_b200.engine import PackedFeatures, packed_distmat

ap = argparse.ArgumentParser()
ap.add_argument("--gallery", type=int, default=1000000)
ap.add_argument("--queries", type=int, default=8192)
ap.add_argument("--reps", type=int, default=5)
args = ap.parse_args()
dev = torch.device("cuda")
peak = 1661.3
if os.path.isfile(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")):
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["bf16_tflops"]
Q, G = args.queries, args.gallery
pitch = (G + 31) // 32 * 32
out = torch.empty((Q, pitch), dtype=torch.float32, device=dev)[:, :G]
print(f"# Q={Q} G={G} cosine bf16 1-pass; peak {peak} TFLOP/s (measured cuBLAS bf16)")
print(f"{'D':>6s} {'ms':>9s} {'TFLOP/s':>9s} {'frac':>6s} {'pack_g ms':>10s}")
for D in (512, 1024, 2048, 2304, 3072, 4096, 6144):
    g = torch.randn(G, D, device=dev, dtype=torch.bfloat16)
    q = torch.randn(Q, D, device=dev, dtype=torch.bfloat16)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); gp = PackedFeatures(g, "cosine", False, "bf16"); b.record(); torch.cuda.synchronize()
    t_pack = a.elapsed_time(b)
    qp = PackedFeatures(q, "cosine", False, "bf16")
    packed_distmat(qp, gp, out); torch.cuda.synchronize()
    ts = []
    for _ in range(args.reps):
        a.record(); packed_distmat(qp, gp, out); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ms = min(ts)
    tf = 2.0 * Q * G * D / ms / 1e9
    print(f"{D:6d} {ms:9.3f} {tf:9.1f} {tf / peak:6.3f} {t_pack:10.3f}")
    del g, q, gp, qp
    torch.cuda.empty_cache()
