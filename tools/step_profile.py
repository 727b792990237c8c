"""Warm per-kernel device times of one adversarial step (torch.profiler / CUPTI), eager and graph replay.
   python tools/step_profile.py [cfg2|cfg5_rank]"""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from pointcloududa_b200.step import AdversarialStep, StepConfig
from torch.profiler import profile, ProfilerActivity

wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
w = dict(bench.WORKLOADS[wl])
dev = torch.device("cuda", 0)
cfg = StepConfig(B=w["B"], C=w["C"], H=w["H"], W=w["W"], N=w["N"], activation=w["activation"], normalize=w["normalize"],
                 return_prob=w["return_prob"])
step = AdversarialStep(cfg, dev, seed=0)
step.load_inputs(bench.make_host_inputs(w, 0, pinned=False), non_blocking=False)
for _ in range(3):
    step.run()
torch.cuda.synchronize()
for mode in ("eager", "graph"):
    if mode == "graph":
        step.capture(warmup=2)
        for _ in range(3):
            step.run()
        torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(5):
            step.run()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    agg = {}
    for e in evs:
        k = e.name[:90]
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1; a[1] += e.device_time if hasattr(e, "device_time") else e.cuda_time
    tot = sum(v[1] for v in agg.values())
    t0 = min(e.time_range.start for e in evs); t1 = max(e.time_range.end for e in evs)
    print(f"== {wl} {mode}: {len(evs)/5:.0f} device events/step, busy {tot/5:.1f} us/step, span {(t1-t0)/5:.1f} us/step")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:32]:
        print(f"  {v[1]/5:8.1f} us {v[0]/5:5.1f} x {v[1]/v[0]:7.2f} us  {k}")
