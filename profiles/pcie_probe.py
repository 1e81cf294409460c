"""Host->device copy rates for the end-to-end path: one big pinned copy vs row chunks on a copy stream."""
import time
import torch

def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    return min(ts) * 1e3

dev = torch.device("cuda")
g = torch.randn(15913, 2304).pin_memory()
q = torch.randn(3368, 2304).pin_memory()
mb = (g.numel() + q.numel()) * 4 / 1e6
gd, qd = torch.empty_like(g, device=dev), torch.empty_like(q, device=dev)
side = torch.cuda.Stream()

def whole():
    qd.copy_(q, non_blocking=True); gd.copy_(g, non_blocking=True)

def chunked(k=4):
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        qd.copy_(q, non_blocking=True)
        step = (15913 + k - 1) // k
        for c0 in range(0, 15913, step):
            gd[c0:c0 + step].copy_(g[c0:c0 + step], non_blocking=True)
    torch.cuda.current_stream().wait_stream(side)

def to_new():
    a = q.to(dev, non_blocking=True); b = g.to(dev, non_blocking=True); return a, b

for name, fn in (("whole, current stream, preallocated", whole), ("4 chunks on a side stream", chunked), ("16 chunks on a side stream", lambda: chunked(16)),
                 (".to(device) fresh allocations", to_new)):
    ms = t(fn)
    print(f"{name:40s} {ms:7.3f} ms  {mb / ms:6.1f} GB/s")
