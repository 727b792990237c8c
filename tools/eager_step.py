"""One eager (no CUDA graph, no concurrent branches) adversarial step between cudaProfilerStart/Stop, for
ncu --profile-from-start off ... python tools/eager_step.py [cfg2|cfg3|cfg5_rank]."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
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
torch.cuda.cudart().cudaProfilerStart()
step.run()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
