"""Run-to-run determinism of the shared MLP forward: which layer's statistics / outputs fluctuate."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from pointcloududa_b200 import ops
from pointcloududa_b200.networks.PointNetCls import shared_mlp
import test_gpu_parity as T

def probe(chans, relus, B, N, pool, precision, mask, runs=40):
    convs, bns, pts = T._tc_stack(chans, relus, B, N, seed=B * 100 + N + 1)
    x = pts.transpose(2, 1)
    ops.tune(3, mask)
    res = []
    with torch.no_grad():
        for it in range(runs):
            for bn in bns: bn.reset_running_stats()
            o = shared_mlp(x, convs, bns, relus, pool=pool, precision=precision).clone()
            res.append([o] + [bn.running_mean.clone() for bn in bns] + [bn.running_var.clone() for bn in bns])
    ops.tune(3, 0)
    L = len(bns)
    worst = [0.0] * (1 + 2 * L)
    nbad = 0
    for r in res[1:]:
        bad = False
        for i, (a, b) in enumerate(zip(r, res[0])):
            d = (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30)
            worst[i] = max(worst[i], d)
            bad = bad or d > 1e-6
        nbad += bad
    print(f"{chans} B={B} N={N} pool={pool} {precision} mask={mask}: runs differing {nbad}/{runs-1}; out {worst[0]:.2e} "
          f"rm {[f'{w:.1e}' for w in worst[1:1+L]]} rv {[f'{w:.1e}' for w in worst[1+L:]]}", flush=True)

for prec, mask in (("fp32", 0), ("bf16", 0), ("bf16", 1)):
    probe([3, 64], [True], 8, 300, False, prec, mask)
    probe([3, 64, 128], [True, True], 8, 300, False, prec, mask)
    probe([3, 64, 128, 1024], [True, True, False], 8, 300, True, prec, mask)
probe([64, 128], [True], 8, 300, False, "bf16", 0)
probe([64, 128], [True], 8, 256, False, "bf16", 0)
probe([128, 1024], [False], 8, 300, True, "bf16", 0)
probe([3, 64, 128, 1024], [True, True, False], 32, 4096, True, "bf16", 0, runs=10)
