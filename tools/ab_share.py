"""A/B of the shared D4 trunk forward (P4 reusing P2's) on the graph-replayed step, L2 flushed.  python tools/ab_share.py [cfg2 ...]"""
import sys, os, statistics
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
import bench
from pointcloududa_b200 import ops
from pointcloududa_b200.step import AdversarialStep, StepConfig
dev = torch.device("cuda:0")
flush = torch.empty(bench.FLUSH_BYTES // 4, device=dev)
for wl in (sys.argv[1:] or ["cfg2"]):
    w = dict(bench.WORKLOADS[wl])
    cfg = StepConfig(B=w["B"], C=w["C"], H=w["H"], W=w["W"], N=w["N"], activation=w["activation"], normalize=w["normalize"],
                     return_prob=w["return_prob"])
    for share in (False, True, False, True):
        step = AdversarialStep(cfg, dev, seed=0, share_trunk=share)
        step.load_inputs(bench.make_host_inputs(w, 0, pinned=False), non_blocking=False)
        n0 = ops.launch_count()
        step.run()
        n1 = ops.launch_count()
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
        print(f"{wl} share_trunk={share}: {n1 - n0} libpcuda launches/step  mean {statistics.mean(ts):.4f} ms  median {statistics.median(ts):.4f}  min {min(ts):.4f}", flush=True)
        step.close(); del step
