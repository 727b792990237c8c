"""Timing experiments on the pooled 128->1024 tensor-core layer (B=32, N=4096), CUPTI device durations over
10 launches each; epilogue variants via pcuda_tune key 4 (see WsParams::dbg).  The dbg variants compute
WRONG results by design (they leave out part of the work)."""
import os, sys
import torch, torch.nn as nn
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointcloududa_b200 import ops
from pointcloududa_b200.networks.PointNetCls import shared_mlp
from torch.profiler import ProfilerActivity, profile
dev = "cuda:0"
B, N = 32, 4096
conv, bn = nn.Conv1d(128, 1024, 1).to(dev), nn.BatchNorm1d(1024).to(dev)
x = torch.rand(B, 128, N, device=dev)
names = {0: "full", 16: "no max/arg-max tracking", 32: "no BN sums", 48: "TMEM loads only (no sums, no max)",
         1: "dbg1: TMEM loads, no math", 2: "no TMEM loads (no epilogue)", 3: "no epilogue, no operand traffic",
         8: "full + cycle report"}
with torch.no_grad():
    for dbg in (0, 16, 32, 48, 1, 2, 3, 8):
        ops.tune(4, dbg)
        for _ in range(3):
            shared_mlp(x, [conv], [bn], [False], pool=True, precision="bf16")
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(10 if dbg != 8 else 1):
                shared_mlp(x, [conv], [bn], [False], pool=True, precision="bf16")
            torch.cuda.synchronize()
        ts = sorted(e.device_time for e in prof.events() if "ws_kernel" in e.name)
        if ts:
            fl = 2.0 * B * N * 128 * 1024
            print(f"dbg={dbg:2d} {names[dbg]:40s}: median {ts[len(ts)//2]:7.2f} us  min {ts[0]:7.2f} us  -> "
                  f"{fl / ts[len(ts)//2] / 1e6:7.1f} TFLOP/s", flush=True)
        else:
            print("dbg", dbg, "no ws_kernel events:", sorted({e.name[:80] for e in prof.events()})[:10])
ops.tune(4, 0)
