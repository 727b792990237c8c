"""Timing experiments on the pooled 128->1024 tensor-core layer (B=32, N=4096): run under
   ncu --metrics gpu__time_duration.sum; epilogue variants via pcuda_tune key 4."""
import os, sys
import torch, torch.nn as nn
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointcloududa_b200 import ops
from pointcloududa_b200.networks.PointNetCls import shared_mlp
dev = "cuda:0"
B, N = 32, 4096
conv, bn = nn.Conv1d(128, 1024, 1).to(dev), nn.BatchNorm1d(1024).to(dev)
x = torch.rand(B, 128, N, device=dev)
with torch.no_grad():
    for dbg in (0, 0):
        ops.tune(4, dbg)
        for _ in range(2):
            shared_mlp(x, [conv], [bn], [False], pool=True, precision="bf16")
        torch.cuda.synchronize()
ops.tune(4, 0)
