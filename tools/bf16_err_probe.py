"""Which piece of the tensor-core backward carries the error against the same-rounding float64 oracle?
Runs tests/test_gpu_step_parity.stack_errors with single pieces switched back to the FP32 kernels
(pcuda_tune(3, mask): 1 forward, 2 pooled dgrad, 4 dense dgrad, 8 wgrad, 16 Gram)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from pointcloududa_b200 import ops
import test_gpu_step_parity as T

cases = [([3, 64, 128, 1024], [True, True, False], 8, 300), ([3, 64, 128, 1024], [True, True, True], 8, 300),
         ([3, 64, 128, 1024], [True, True, False], 4, 1024)]
if len(sys.argv) > 1:
    cases.append(([3, 64, 128, 1024], [True, True, False], 32, 4096))
for chans, relus, B, N in cases:
    for mask in (0, 30, 2, 4, 8, 16, 28, 26, 22, 14):
        ops.tune(3, mask)
        try:
            e = T.stack_errors(chans, relus, B, N, seed=1000 + 7 * B + N)
        finally:
            ops.tune(3, 0)
        e.pop("masked_pairs")
        print(f"B={B} N={N} relu_last={relus[-1]} fp32-mask={mask:2d}: " + " ".join(f"{k}={m:.1e}/{n:.1e}" for k, (m, n) in e.items() if k != "out"), flush=True)
