"""A/B of the parameter L2 prefetch on the graph-replayed step (L2 flushed before every step)."""
import sys, os, statistics
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from bench import WORKLOADS, build_step, FLUSH_BYTES
dev = torch.device("cuda:0")
flush = torch.empty(FLUSH_BYTES // 4, device=dev)
for wl in (sys.argv[1:] or ["cfg2"]):
    for pf in (False, True, False, True):
        step, hf, rh = build_step(dict(WORKLOADS[wl]), 0, dev, "bf16", True)
        step.prefetch = pf
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
        print(f"{wl} prefetch={pf}: mean {statistics.mean(ts):.4f} ms  median {statistics.median(ts):.4f}  min {min(ts):.4f}", flush=True)
        step.close(); del step
