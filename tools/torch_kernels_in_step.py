"""Which framework (at::) kernels does one eager adversarial step launch, and from which line of step.py / ops.py?
python tools/torch_kernels_in_step.py [cfg2]"""
import sys, os, collections
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from torch.profiler import profile, ProfilerActivity
import bench
from pointcloududa_b200.step import AdversarialStep, StepConfig
dev = torch.device("cuda:0")
w = dict(bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg2"])
cfg = StepConfig(B=w["B"], C=w["C"], H=w["H"], W=w["W"], N=w["N"], activation=w["activation"], normalize=w["normalize"],
                 return_prob=w["return_prob"])
step = AdversarialStep(cfg, dev, seed=0, concurrent=False)
step.load_inputs(bench.make_host_inputs(w, 0, pinned=False), non_blocking=False)
for _ in range(3):
    step.run()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True, record_shapes=True) as prof:
    step.run()
    torch.cuda.synchronize()
n_all = 0
rows = collections.Counter()
for e in prof.events():
    if e.device_type != torch.autograd.DeviceType.CPU or not e.kernels:
        continue
    for k in e.kernels:
        n_all += 1
        if "pcuda" in k.name:
            continue
        stack = [s for s in (e.stack or []) if "pointcloududa_b200" in s or "bench.py" in s]
        where = stack[0].strip() if stack else "(autograd engine / no python frame)"
        rows[(e.name, k.name.split("<")[0][:60], str(e.input_shapes)[:70], where[-90:])] += 1
print(f"{n_all} device launches in the step; framework kernels:")
for (op, kn, shp, where), n in sorted(rows.items(), key=lambda kv: -kv[1]):
    print(f"{n:3d} x {op:28s} {kn:50s} {shp:70s} {where}")
