"""Approximate critical path of the graph-replayed adversarial step from CUPTI kernel records: walk back from the
kernel that ends last, each time to the kernel whose end is the latest one not after the current start (+0.3 us).
Prints the chain with durations and the gaps between links.   python tools/critical_path.py [cfg2|cfg5_rank]
The step is captured WITHOUT programmatic dependent launch (pcuda_tune(8, 1)): with it a kernel's record starts when the
kernel becomes resident, i.e. includes the time it waits for its predecessor (see tools/stream_chains.py for that view).
The profiler itself stretches the step (branches overlap less than in an unprofiled replay): read shares, not totals."""
import os, sys
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
from pointcloududa_b200 import ops
ops.tune(8, 1)
step = AdversarialStep(cfg, dev, seed=0)
step.load_inputs(bench.make_host_inputs(w, 0, pinned=False), non_blocking=False)
for _ in range(3):
    step.run()
step.capture(warmup=2)
for _ in range(5):
    step.run()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step.run()
    torch.cuda.synchronize()
evs = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA),
             key=lambda e: e.time_range.start)
t0 = evs[0].time_range.start
last = max(evs, key=lambda e: e.time_range.end)
print(f"{wl}: {len(evs)} device events, span {last.time_range.end - t0:.1f} us, busy {sum(e.time_range.end - e.time_range.start for e in evs):.1f} us")
chain, cur = [], last
while cur is not None:
    chain.append(cur)
    cands = [e for e in evs if e is not cur and e.time_range.end <= cur.time_range.start + 0.3]
    cur = max(cands, key=lambda e: e.time_range.end) if cands else None
chain.reverse()
short = lambda n: n.replace("(anonymous namespace)::", "").replace("void ", "").replace("pcuda::", "").split("(")[0][:58]
tk = tg = 0.0
prev_end = t0
agg = {}
for e in chain:
    gap = e.time_range.start - prev_end
    dur = e.time_range.end - e.time_range.start
    tk += dur; tg += max(gap, 0.0)
    a = agg.setdefault(short(e.name), [0, 0.0]); a[0] += 1; a[1] += dur
    print(f"  +{e.time_range.start - t0:8.1f} us  gap {gap:6.1f}  dur {dur:6.1f}  {short(e.name)}")
    prev_end = e.time_range.end
print(f"chain: {len(chain)} kernels, kernel time {tk:.1f} us, gaps {tg:.1f} us")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:20]:
    print(f"  {v[1]:7.1f} us {v[0]:3d} x  {k}")
