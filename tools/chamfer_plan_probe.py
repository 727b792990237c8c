"""Chamfer forward: rows-per-thread / CTA-size variants of the prefiltered kernel over (B, N) — data for make_plan."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointcloududa_b200 import ops
from pointcloududa_b200.synth import synth_clouds
DEV = "cuda:0"
flush = torch.empty(256 * 1024 * 1024 // 4, device=DEV)

def timeit(fn, iters):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.fill_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]

sizes = [(8, 300), (4, 1024), (8, 1024), (16, 1024), (32, 1024), (64, 1024), (8, 2048), (16, 2048), (32, 2048), (64, 2048),
         (4, 4096), (8, 4096), (16, 4096), (32, 4096), (64, 4096), (8, 8192), (16, 8192), (64, 8192), (8, 16384), (64, 16384), (64, 256), (64, 512)]
for B, N in sizes:
    x, y = synth_clouds(B, N, seed=5, mode="near")
    x, y = x.to(DEV), y.to(DEV)
    row = []
    for var in (0, 1064, 1128, 2128, 4128, 8128):
        ops.tune(1, var)
        try:
            ms = timeit(lambda: ops.chamfer_nn(x, y), 7)
        finally:
            ops.tune(1, 0)
        row.append((var, ms))
    best = min(row[1:], key=lambda r: r[1])
    print(f"B{B:3d} N{N:6d}: " + "  ".join(f"{v}:{ms*1e3:8.1f}us" for v, ms in row) + f"   best {best[0]} ({row[0][1] / best[1]:.2f}x of default)", flush=True)
