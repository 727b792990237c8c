"""Launch each hot kernel a few times at roofline-sized shapes, for ncu captures."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointcloududa_b200 import ops  # noqa: E402
from pointcloududa_b200.synth import synth_clouds  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "all"
dev = "cuda:0"
if which in ("all", "entropy"):
    for fast in (0,):
        ops.tune(0, fast)
        z = torch.randn(256, 5, 256, 256, device=dev) * 3
        g = torch.randn_like(z)
        for _ in range(2):
            zr = z.requires_grad_(True)
            m = ops.entropy_map(zr, "softmax", True)
            torch.autograd.grad(m, zr, g)
        zs = torch.randn(256, 4, 224, 224, device=dev) * 3
        for _ in range(2):
            zr = zs.requires_grad_(True)
            m = ops.entropy_map(zr, "sigmoid", False)
            torch.autograd.grad(m, zr, torch.ones_like(m))
        del z, g, zs
    ops.tune(0, 0)
if which in ("all", "chamfer"):
    for B, N in ((64, 4096), (64, 16384)):
        x, y = synth_clouds(B, N, seed=5, mode="near")
        x, y = x.to(dev), y.to(dev)
        for var in (0,):
            ops.tune(1, var)
            for _ in range(2):
                ops.chamfer_nn(x, y)
        ops.tune(1, 0)
if which in ("all", "mlp"):
    from pointcloududa_b200.networks.PointNetCls import PointNetCls
    net = PointNetCls().to(dev).train()
    pts = torch.rand(32, 4096, 3, device=dev, requires_grad=True)
    for _ in range(2):
        o = net(pts.transpose(2, 1))[0]
        o.sum().backward()
if which == "mlp_small":
    # cfg2-sized D4 pass (B=8, N=300): the latency-bound helper kernels
    from pointcloududa_b200.networks.PointNetCls import PointNetCls
    net = PointNetCls().to(dev).train()
    pts = torch.rand(8, 300, 3, device=dev, requires_grad=True)
    for _ in range(3):
        o = net(pts.transpose(2, 1))[0]
        o.sum().backward()
torch.cuda.synchronize()
