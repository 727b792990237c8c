"""A/B of one pcuda_tune key on the graph-replayed step (L2 flushed before every step, 40 steps, settings interleaved
twice).   python tools/ab_tune.py KEY VALUE[,VALUE...] [workload ...]      e.g.  ab_tune.py 11 0,1 cfg2 cfg3"""
import sys, os, statistics
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from bench import WORKLOADS, build_step, FLUSH_BYTES
from pointcloududa_b200 import ops
key = int(sys.argv[1])
values = [int(v) for v in sys.argv[2].split(",")]
dev = torch.device("cuda:0")
flush = torch.empty(FLUSH_BYTES // 4, device=dev)
for wl in (sys.argv[3:] or ["cfg2"]):
    for v in values * 2:
        ops.tune(key, v)
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
        print(f"{wl} tune({key}, {v}): mean {statistics.mean(ts):.4f} ms  median {statistics.median(ts):.4f}  min {min(ts):.4f}", flush=True)
        step.close(); del step
ops.tune(key, 0)
