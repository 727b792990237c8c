"""A/B of the pooled dgrad's sorted sparse-row kernel (tuning key 10: -1 = the one-warp-per-point kernel) on the graph-replayed
step, L2 flushed; also prints the kernel's own durations from a profiler pass."""
import sys, os, statistics
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from torch.profiler import profile, ProfilerActivity
from bench import WORKLOADS, build_step, FLUSH_BYTES
from pointcloududa_b200 import ops
dev = torch.device("cuda:0")
flush = torch.empty(FLUSH_BYTES // 4, device=dev)
for wl in (sys.argv[1:] or ["cfg2"]):
    for heavy in (-1, 0, -1, 0, -1, 0):
        ops.tune(10, heavy)
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
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            flush.fill_(1.0); step.run(); torch.cuda.synchronize()
        sp = [e.time_range.end - e.time_range.start for e in prof.events()
              if e.device_type == torch.autograd.DeviceType.CUDA and "pool_sparse" in e.name]
        print(f"{wl} sorted_key={heavy}: mean {statistics.mean(ts):.4f} ms  median {statistics.median(ts):.4f}  min {min(ts):.4f}"
              f"   pool_sparse us: {' '.join(f'{t:.1f}' for t in sp)}", flush=True)
        step.close(); del step
ops.tune(10, 0)
