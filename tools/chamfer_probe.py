"""Chamfer forward: kernel variants x cfg-4 sizes (A/B for make_plan). Writes gpurun_out/chamfer_probe.json."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointcloududa_b200 import ops  # noqa: E402
from pointcloududa_b200.synth import synth_clouds  # noqa: E402

DEV = "cuda:0"


def timeit(fn, iters, flush):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.fill_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    flush = torch.empty(256 * 1024 * 1024 // 4, device=DEV)
    out = {}
    sizes = [(64, 256), (64, 512), (64, 1024), (64, 2048), (64, 4096), (64, 8192), (64, 16384), (8, 300), (32, 1024), (32, 4096)]
    if len(sys.argv) > 1 and sys.argv[1] == "quick":
        sizes = [(64, 2048), (64, 4096), (64, 16384), (32, 4096)]
    for mode in ("near", "uniform"):
        for B, N in sizes:
            if mode == "uniform" and (B, N) not in ((64, 4096), (64, 16384), (64, 1024)):
                continue
            x, y = synth_clouds(B, N, seed=5, mode=mode)
            x, y = x.to(DEV), y.to(DEV)
            pairs = 2.0 * B * N * N
            variants = (0, 1064, 1128, 2128, 4128, 8128, 200000, 204128, 208128)
            if len(sys.argv) > 1 and sys.argv[1] == "quick":
                variants = (0, 4128, 304128, 204128, (4128, 8), (4128, 32), (4128, -1))
            for var in variants:
                if N >= 8192 and var in (1064, 1128):
                    continue
                seed = 0
                if isinstance(var, tuple):
                    var, seed = var
                ops.tune(1, var)
                ops.tune(5, seed)
                try:
                    ms = timeit(lambda: ops.chamfer_nn(x, y), 5 if N > 4096 else 10, flush)
                finally:
                    ops.tune(1, 0)
                    ops.tune(5, 0)
                if seed:
                    var = f"{var}_seed{seed}"
                out[f"{mode}/B{B}_N{N}/var{var}"] = dict(ms=ms, tpairs_s=pairs / ms / 1e9)
                print(f"{mode} B{B} N{N} var{var}: {ms:.4f} ms  {pairs / ms / 1e9:.3f} Tpairs/s", flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/chamfer_probe.json", "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
