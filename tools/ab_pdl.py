"""A/B of programmatic dependent launch on the cfg2 step (graph replay, L2 flushed): off, on the launches of kernels with a
prologue before their grid-dependency wait (the default), on every launch."""
import sys, os, statistics
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from bench import WORKLOADS, build_step, FLUSH_BYTES
from pointcloududa_b200 import ops
dev = torch.device("cuda:0")
flush = torch.empty(FLUSH_BYTES // 4, device=dev)
for wl in (sys.argv[1:] or ["cfg2"]):
    for no_pdl in (1, 4, 0, 1, 4, 0):
        ops.tune(8, no_pdl)
        step, hf, rh = build_step(dict(WORKLOADS[wl]), 0, dev, "bf16", True)
        step.capture(warmup=2)
        for _ in range(5):
            step.run()
        torch.cuda.synchronize()
        ts = []
        for _ in range(40):
            flush.fill_(1.0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); step.run(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        name = {1: "off", 4: "prologue-kernels", 0: "every-launch", 3: "two-waves"}[no_pdl]
        print(f"{wl} pdl={name}: mean {statistics.mean(ts):.4f} ms  median {statistics.median(ts):.4f}  min {min(ts):.4f}", flush=True)
        step.close(); del step
ops.tune(8, 0)
