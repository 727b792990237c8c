"""One-shot GPU probe: kernel timings for A/B decisions (not the bench). Writes gpurun_out/probe.json."""
import json
import math
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointcloududa_b200 import ops  # noqa: E402
from pointcloududa_b200.synth import synth_clouds, synth_logits  # noqa: E402

DEV = "cuda:0"
out = {}


def timeit(fn, iters=20, warm=3, flush=None):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.fill_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    torch.cuda.init()
    flush = torch.empty(256 * 1024 * 1024 // 4, device=DEV)
    # ---- entropy: cfg-5 per-rank shard (B=32,C=5,256^2 = 42 MB)... and the full 336 MB tensor
    for name, shape, act, norm in [("cfg5_full", (256, 5, 256, 256), "softmax", True),
                                   ("cfg5_rank", (32, 5, 256, 256), "softmax", True),
                                   ("cfg1", (8, 4, 224, 224), "sigmoid", False),
                                   ("sig_big", (256, 4, 224, 224), "sigmoid", False)]:
        z = torch.randn(shape, device=DEV) * 3
        nbytes = z.numel() * 4
        g = torch.randn_like(z)
        for fast in (0, 1):
            ops.tune(0, fast)
            m = torch.empty_like(z)
            med, best = timeit(lambda: ops.entropy_map(z, act, norm), flush=flush)
            out[f"entropy_fwd/{name}/fast{fast}"] = dict(ms=med, best_ms=best, gbs=2 * nbytes / med / 1e6)
            med, best = timeit(lambda: ops.entropy_map(z, act, norm, return_prob=True, return_mean=True), flush=flush)
            out[f"entropy_fwd_p_mean/{name}/fast{fast}"] = dict(ms=med, best_ms=best, gbs=3 * nbytes / med / 1e6)
            zr = z.clone().requires_grad_(True)
            mm = ops.entropy_map(zr, act, norm)
            med, best = timeit(lambda: torch.autograd.grad(mm, zr, g, retain_graph=True), flush=flush)
            out[f"entropy_bwd/{name}/fast{fast}"] = dict(ms=med, best_ms=best, gbs=3 * nbytes / med / 1e6)
        ops.tune(0, 0)
        del z, g, zr, mm
    # torch eager reference for the map, for context
    z = torch.randn(32, 5, 256, 256, device=DEV) * 3
    def eager():
        p = torch.softmax(z, 1)
        return -1.0 * p * torch.log(p + 1e-7) / math.log(5)
    med, best = timeit(eager, flush=flush)
    out["entropy_fwd/cfg5_rank/torch_eager"] = dict(ms=med, best_ms=best)
    del z
    # ---- chamfer variants
    for B, N in [(64, 4096), (64, 16384), (32, 1024), (8, 300)]:
        x, y = synth_clouds(B, N, seed=5, mode="near")
        x, y = x.to(DEV), y.to(DEV)
        pairs = 2.0 * B * N * N
        for var in (0, 1128, 2128, 4128, 8128, 104128, 102128):
            if N <= 300 and var not in (0, 1064, 1128):
                continue
            ops.tune(1, var)
            med, best = timeit(lambda: ops.chamfer_nn(x, y), iters=10 if N > 4096 else 20)
            out[f"chamfer_fwd/B{B}_N{N}/var{var}"] = dict(ms=med, best_ms=best, tpairs_s=pairs / med / 1e9)
        ops.tune(1, 0)
        xr = x.clone().requires_grad_(True)
        loss = ops.chamfer_nn(xr, y)[0]
        med, best = timeit(lambda: torch.autograd.grad(loss, xr, retain_graph=True))
        out[f"chamfer_bwd/B{B}_N{N}"] = dict(ms=med, best_ms=best)
    # ---- point MLP (fp32 path)
    from pointcloududa_b200.networks.PointNetCls import PointNetCls
    for B, N in [(8, 300), (32, 1024), (32, 4096)]:
        for prec in ("fp32", "bf16"):
            net = PointNetCls(precision=prec).to(DEV).train()
            pts = torch.rand(B, N, 3, device=DEV)
            def fwd():
                return net(pts.transpose(2, 1))[0]
            med, best = timeit(fwd, iters=10)
            flops = 2 * 278912 * B * N
            out[f"pointnet_fwd/B{B}_N{N}/{prec}"] = dict(ms=med, best_ms=best, tflops=flops / med / 1e9)
            def fwdbwd():
                o = net(pts.transpose(2, 1))[0]
                o.sum().backward()
            med, best = timeit(fwdbwd, iters=10)
            out[f"pointnet_fwdbwd/B{B}_N{N}/{prec}"] = dict(ms=med, best_ms=best)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/probe.json", "w") as f:
        json.dump(out, f, indent=1)
    for k, v in out.items():
        print(k, {a: round(b, 4) for a, b in v.items()})


if __name__ == "__main__":
    main()
