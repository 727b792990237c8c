#!/usr/bin/env python
"""bench.py — adversarial-step throughput of the PointCloudUDA hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2]

One "step" = the hot-path work of one train_epoch iteration (pointcloududa_b200/step.py):
entropy map fwd+bwd / fwd, Chamfer fwd+bwd / fwd, D4 (PointNetCls) 3x fwd + 3x bwd, BCE, D4 gradient
all-reduce (N > 1) and SGD step — on synthetic slices / clouds of the reference's shapes.
Prints ONE JSON line (rank 0).  See DESIGN.md §Measurement for the definition of every field.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "uda_adversarial_steps_per_s"
UNIT = "steps/s"

WORKLOADS = {
    # BASELINE.json configs[1]: MS-CMRSeg step on 1xB200 (per-rank batch under weak scaling)
    "cfg2": dict(B=8, C=4, H=224, W=224, N=300, activation="sigmoid", normalize=False, return_prob=False),
    # configs[2]: MM-WHS step, global batch 32 (strong scaling: 32 / n_gpus clouds per rank, 4 per rank on 8 GPUs)
    "cfg3": dict(B=32, C=5, H=256, W=256, N=1024, activation="softmax", normalize=True, return_prob=True),
    # configs[4] per-rank shard (32 of 256 samples)
    "cfg5_rank": dict(B=32, C=5, H=256, W=256, N=4096, activation="softmax", normalize=True, return_prob=True),
}


# L2 flush between timed steps: a write larger than the 126 MB L2
FLUSH_BYTES = 160 * 1024 * 1024


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm=d["hbm_gbs"], tensor=d["bf16_tflops"], tensor_sustained=d.get("bf16_tflops_sustained"),
                    sm_max_mhz=d.get("sm_max_mhz", 1965.0), source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tensor=1590.0, tensor_sustained=1400.0, source="fallback (B200_PROFILING.md)")


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
def make_host_inputs(w: dict, rank: int, pinned: bool):
    import torch
    from pointcloududa_b200.synth import synth_clouds, synth_logits
    B, C, H, W, N = w["B"], w["C"], w["H"], w["W"], w["N"]
    seed = 1234 + 97 * rank
    xS, yA = synth_clouds(B, N, seed=seed + 1, mode="near")
    xT, yB = synth_clouds(B, N, seed=seed + 2, mode="uniform")
    host = {
        "oS": synth_logits(B, C, H, W, seed=seed + 3), "oT": synth_logits(B, C, H, W, seed=seed + 4),
        "grad_mapT": synth_logits(B, C, H, W, seed=seed + 5, std=1.0),
        "vertS": xS, "vertT": xT, "vertexA": yA, "vertexB": yB,
    }
    if pinned:
        host = {k: v.pin_memory() for k, v in host.items()}
    return host


def cpu_step(w: dict, host_np: dict, params, spec):
    """The same step on the host cores with the oracle port (reference restatement)."""
    import numpy as np
    import oracle
    act, norm = w["activation"], w["normalize"]
    # P1
    o = oracle.chamfer_fwd(host_np["vertS"], host_np["vertexA"])
    oracle.chamfer_bwd64(host_np["vertS"], host_np["vertexA"], o["d1"], o["i1"], o["d2"], o["i2"])
    # P2
    oracle.entropy_fwd(host_np["oT"], act, norm, want_p=w["return_prob"], want_mean=True)
    oracle.entropy_bwd64(host_np["oT"], host_np["grad_mapT"], None, 0.0, act, norm)
    oracle.chamfer_fwd(host_np["vertT"], host_np["vertexB"])
    losses = []
    for pts in (host_np["vertT"], host_np["vertS"], host_np["vertT"]):
        x = np.ascontiguousarray(pts.transpose(0, 2, 1))
        logit, _, _, cache = oracle.pointnet_forward(params, x, spec)
        dlogit = ((1.0 / (1.0 + np.exp(-logit))) - 1.0) / logit.shape[0]
        oracle.pointnet_backward(dlogit.astype(np.float32), cache)
        losses.append(float(logit.mean()))
    # P3 entropy of the source logits (forward only)
    oracle.entropy_fwd(host_np["oS"], act, norm, want_p=w["return_prob"])
    return losses


def workload_label(name: str, w: dict) -> str:
    kind = {"cfg2": "MS-CMRSeg", "cfg3": "MM-WHS", "cfg3_rank": "MM-WHS (cfg-3 per-rank shard)",
            "cfg5_rank": "large-batch (cfg-5 per-rank shard)"}.get(name, name)
    return (f"{name}: {kind} adversarial step, per-rank batch {w['B']}, {w['C']}x{w['H']}x{w['W']} logits, "
            f"{w['N']}-pt clouds")


_T0 = time.time()


def progress(msg: str) -> None:
    """Phase marker on stderr (rank-tagged): where a multi-rank run spends its time, and where it stops if it ever hangs."""
    print(f"[bench rank {os.environ.get('RANK', '0')} +{time.time() - _T0:6.1f}s] {msg}", file=sys.stderr, flush=True)


def bench_config(name: str, w: dict, world: int, precision: str = "bf16", use_graph: bool = True) -> dict:
    """`config` of the JSON line — the SAME dict for the GPU arm and for `--impl reference` (the driver compares
    them), so the statements about L2 / graphs name the arm they apply to."""
    return {"workload": workload_label(name, w), "per_rank": w, "global_batch": w["B"] * world,
            "parallelism": f"dp{world} (batch-sharded, D4 grad all-reduce)",
            "l2": f"GPU arm: flushed before every step ({FLUSH_BYTES >> 20} MiB write, L2 = 126 MB): outside the per-step "
                  "CUDA-event intervals of `value`, inside the single interval of `e2e`; reference arm: host cores, "
                  "working set larger than the last-level cache slice of a core",
            "e2e_pipeline": "GPU arm: copy stream uploads step i+1 (one flat pinned buffer) into a staging set while step i "
                            "runs; one device-to-device move into the step's static inputs",
            "cuda_graph": f"GPU arm: {bool(use_graph)}", "mlp_precision": f"GPU arm: {precision}; reference arm: fp32",
            "d4_passes": "three full D4 forward + backward passes per step in both arms (P2, P3, P4 each compute their own "
                         "forward); the GPU arm's `shared_trunk` object is the same step with P4 reusing the part of P2's "
                         "forward that is identical (same cloud, same weights), bit-identical results"}


def reference_available() -> bool:
    try:
        from baseline import ref_step
        return ref_step.available()
    except Exception:
        return False


def run_cpu_baseline(w: dict, budget_s: float, min_steps: int = 2, max_steps: int = 50, warmup: int = 1):
    """The reference path on the host cores.  kind="reference": the reference's own PyTorch code (baseline/_ref,
    torch eager, every host core); kind="port" (only when baseline/_ref is not installed): the oracle port."""
    cores = os.cpu_count() or 1
    if reference_available():
        import torch
        from baseline import ref_step
        torch.set_num_threads(cores)
        host = make_host_inputs(w, 0, pinned=False)
        ref = ref_step.ReferenceStep(w, "cpu", seed=0)
        step_fn = lambda: ref.run(host)
        kind = "reference"
        what = ("the reference's own code (utils/loss.py batch_NN_loss with the one-token in-memory device patch, "
                "networks/PointNetCls.py, inline entropy expressions) under torch eager")
    else:
        import oracle
        host = make_host_inputs(w, 0, pinned=False)
        host_np = {k: v.numpy() for k, v in host.items()}
        spec = oracle.PointNetSpec()
        params = oracle.make_params(spec, 0)
        oracle.set_num_threads(cores)
        step_fn = lambda: cpu_step(w, host_np, params, spec)
        kind = "port"
        what = "oracle port: C+OpenMP entropy/Chamfer, numpy/BLAS PointNet (baseline/_ref not installed)"
    for _ in range(max(1, warmup)):
        step_fn()
    ts = []
    t_start = time.perf_counter()
    while len(ts) < max_steps and (len(ts) < min_steps or time.perf_counter() - t_start < budget_s):
        t0 = time.perf_counter()
        step_fn()
        ts.append(time.perf_counter() - t0)
    med = statistics.median(ts)
    return {"value": 1.0 / med, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": f"{len(ts)} full steps of the same workload ({what}), median",
            "ms_per_step": med * 1e3, "ms_per_step_mean": statistics.mean(ts) * 1e3}


def run_gpu_eager_baseline(w: dict, dev, host: dict, flush, steps: int = 20, warmup: int = 3):
    """The UNPATCHED reference (baseline/_ref) executed under torch eager on this same B200 with its stock
    cuBLAS / cuDNN / ATen kernels: the bar the sm_100a kernels have to beat.  Inputs resident,
    L2 flushed between steps, per-step CUDA events like `value`."""
    import torch
    if not reference_available():
        return {"unavailable": "baseline/_ref not installed (python baseline/make_ref.py needs /root/reference)"}
    from baseline import ref_step
    ref = ref_step.ReferenceStep(w, dev, seed=0)
    I = {k: v.to(dev) for k, v in host.items()}
    for _ in range(warmup):
        ref.run(I)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for e0, e1 in ev:
        flush.fill_(1.0)
        e0.record(); ref.run(I); e1.record()
    torch.cuda.synchronize()
    ts = [e0.elapsed_time(e1) for e0, e1 in ev]
    ms = statistics.mean(ts)
    del ref, I
    torch.cuda.empty_cache()
    return {"value": 1e3 / ms, "unit": UNIT, "ms_per_step": ms, "ms_per_step_min": min(ts), "steps": steps,
            "what": "unmodified reference code (baseline/_ref: batch_NN_loss, PointNetCls, inline entropy expressions, "
                    "torch.optim.SGD) under torch eager on this GPU, fp32 (TF32 off), inputs resident, CUDA events"}


# ------------------------------------------------------------------------------------------------
def _splits(w):
    m = w["B"] * w["N"]
    return max(1, min(2 * 148, (m + 127) // 128))


def _M(w):
    return float(w["B"]) * w["N"]


KERNEL_WORK = {
    # kernel-name fragment -> (algorithmic FLOPs per launch or None, compulsory HBM bytes per launch: every operand once,
    # every result once) as functions of the per-rank workload.  Default D4 (3->64->128->1024 twice: STN3d trunk and
    # main trunk), M = B*N points, C = 1024, K = 128; where one kernel serves several layer shapes the widest is taken
    # (an upper bound of the work).  The roofline of a kernel is the LARGER of flops / tensor peak and bytes / HBM peak:
    # only the pooled 128->1024 GEMM (1024 FLOP per byte of operand) is tensor-bound, the narrow GEMMs move a bf16
    # operand and an fp32 result per ~20 FLOP/B and sit on the HBM roof.
    "ws_kernel<1>": (lambda w: 2.0 * _M(w) * 128 * 1024, lambda w: _M(w) * 128 * 2.0 + 1024 * 128 * 2.0 + w["B"] * 1024 * 8.0),
    "ws_kernel<0>": (lambda w: 2.0 * _M(w) * 64 * 128, lambda w: _M(w) * (64 * 2.0 + 128 * 4.0)),            # a bf16 in, y fp32 out
    "ws_kernel<2>": (lambda w: 2.0 * _M(w) * 128 * 64, lambda w: _M(w) * (128 * 2.0 + 64 * 4.0 + 64 * 4.0)),   # dy bf16, y_prev in, dz out
    "ws_kernel<3>": (lambda w: 2.0 * _M(w) * 128 * 128, lambda w: _M(w) * (128 * 2.0 + 128 * 4.0 + 128 * 4.0)),  # a bf16, y_prev in, dz out
    "pt_kernel": (lambda w: 2.0 * _M(w) * 128 * 128, lambda w: _M(w) * (128 + 128) * 2.0),                   # both operands bf16 slabs
    "entropy_fwd": (None, lambda w: (8.0 + (4.0 if w["return_prob"] else 0.0)) * w["B"] * w["C"] * w["H"] * w["W"]),
    "entropy_bwd": (None, lambda w: 12.0 * w["B"] * w["C"] * w["H"] * w["W"]),
    "chamfer_nn": (lambda w: 3.0 * 2.0 * w["B"] * w["N"] * w["N"], lambda w: 48.0 * w["B"] * w["N"]),          # 3 FP32 lane-ops per ordered pair (prefilter)
    "chamfer_bwd": (None, lambda w: 36.0 * 2 * w["B"] * w["N"]),
    "pack_act_kernel": (None, lambda w: (4.0 + 2.0) * _M(w) * 128),    # fp32 in, bf16 out (widest layer)
    "pack_dy_kernel": (None, lambda w: (8.0 + 2.0) * _M(w) * 128),     # dz and y in, bf16 out
    "pack_w_kernel": (None, lambda w: (4.0 + 2.0) * 1024 * 128),
    # sparse rows of the pooled-layer dgrad: arg/coef, one W row per (cloud, channel), read-modify-write of the
    # <= min(C, N) selected points per cloud (y_prev, dz_prev in, dz_prev out), K = 128, C = 1024
    "pool_sparse": (None, lambda w: w["B"] * 1024 * (8.0 + 4.0 * 128) + w["B"] * min(1024, w["N"]) * 128 * 12.0),
    "pool_dw_kernel": (None, lambda w: 1024 * 128 * 8.0 + 128 * 128 * 4.0 + w["B"] * 1024 * (8.0 + 4.0 * 128)),
    "pool_sel_kernel": (None, lambda w: w["B"] * 1024 * (4.0 * 128 + 24.0) + 1024 * 128 * 4.0),   # selected rows + W + out/arg/grad/coef
    "pool_coef_kernel": (None, lambda w: w["B"] * 1024 * 8.0 + 1024 * 28.0),
    "pool_q_kernel": (None, lambda w: 1024 * 128 * 4.0 + 1024 * 8.0 + 8 * 128 * (128 + 1) * 4.0),   # W, kappa, alpha in, split partials out
    "q_finish_kernel": (None, lambda w: 8 * 128 * (128 + 1) * 4.0 + 128 * 128 * 6.0),                # partials in, Q fp32 + bf16 out
    "gram_finish_kernel": (None, lambda w: (_splits(w) + 1) * 128 * 128 * 4.0 + _splits(w) * 128 * 8.0),
    "pool_u_kernel": (None, lambda w: 1024 * 128 * 4.0 + 128 * 128 * 4.0 + 1024 * 4.0 + 1024.0),
    "pool_finalize_kernel": (None, lambda w: w["B"] * 1024 * 16.0),
    "act_colsum_kernel": (None, lambda w: 4.0 * _M(w) * 128),
    "abar_kernel": (None, lambda w: _splits(w) * 128 * 8.0),
    "gram_center_kernel": (None, lambda w: 128 * 128 * 12.0),
    "reduce_partials_kernel": (None, lambda w: (_splits(w) + 1) * 1024 * 128 * 4.0),              # widest: dW of the pooled layer
    "reduce_partials_split_kernel": (None, lambda w: (148 + 1) * 64 * 4 * 4.0),                    # one partial per SM of the 3 -> 64 layer's dW
    "mlp_fwd_narrow_kernel": (None, lambda w: (12.0 + 4.0 * 64) * _M(w)),               # 3 -> 64: cloud in, y out
    "wgrad_narrow_kernel": (None, lambda w: (12.0 + 8.0 * 64) * _M(w)),
    "dgrad_input_narrow_kernel": (None, lambda w: (8.0 * 64 + 12.0) * _M(w)),
    "input_transform_bwd_kernel": (None, lambda w: 36.0 * _M(w)),
    "point_transform_kernel": (None, lambda w: 2 * 4.0 * 64 * _M(w)),
    "bn_finalize_pivot_kernel": (None, lambda w: 1024 * 40.0),
    "bn_bwd_coef_kernel": (None, lambda w: 1024 * 44.0),
    "bn_running_update": (None, lambda w: 3968 * 3 * 16.0),
    "fc_fwd_kernel": (None, lambda w: 4.0 * (1024 * 512 + 512 * 256 + 256 * 9) / 3 + 4.0 * w["B"] * (1024 + 512)),   # mean layer of a head
    "fc_bwd_kernel": (None, lambda w: 2 * 4.0 * (1024 * 512 + 512 * 256 + 256 * 9) / 3 + 4.0 * w["B"] * (1024 + 512)),
    "fc_pull_kernel": (None, lambda w: 4.0 * 512 * 1024 + 4.0 * w["B"] * (1024 + 512)),
    "fc_head": (None, lambda w: 4.0 * (1024 * 512 + 512 * 256 + 256 * 9) + 4.0 * w["B"] * (1024 + 512 + 256)),
    "sgd_momentum_sum": (None, lambda w: 1.6e6 * 28.0),                        # two gradient lists, momentum, param in / bucket, momentum, param out
    "sgd": (None, lambda w: 1.6e6 * 16.0),                                      # grad, momentum in / momentum, param in+out
    "grad_sum_pack_kernel": (None, lambda w: 1.6e6 * 12.0),                    # two gradient lists in, the bucket out
    "p2p_allreduce_kernel": (None, lambda w: 1.6e6 * 8.0),                     # per GPU: (R-1)/R of the bucket read + written over NVLink
    "bce_logits_kernel": (None, lambda w: w["B"] * 8.0 + 8.0),
}


def kernel_roofline(name: str, w: dict, seconds: float, pk: dict, traffic=None) -> dict:
    """Roofline record of one launch of kernel `name` that took `seconds`: achieved rate against the roof that bounds
    it — tensor peak if flops / peak_tensor > bytes / peak_hbm, else HBM — with both fractions kept for reference
    ("latency": the launch's whole roofline time is below 1 us, so neither roof can be approached at this size).
    When the ncu DRAM traffic of the launch is below half of its algorithmic bytes the working set is L2-resident and
    the launch is a chain of dependent L2 round trips: labelled "l2-latency" (the HBM fraction is then not a statement
    about the kernel)."""
    for frag, (flops_f, bytes_f) in KERNEL_WORK.items():
        if frag not in name:
            continue
        nbytes = bytes_f(w)
        flops = flops_f(w) if flops_f is not None else None
        rec = {"algorithmic_bytes_per_launch": nbytes, "hbm_gbs": nbytes / seconds / 1e9,
               "frac_hbm": nbytes / seconds / 1e9 / pk["hbm"]}
        is_chamfer = "chamfer_nn" in name
        if flops is not None and not is_chamfer:
            rec.update({"algorithmic_flops_per_launch": flops, "tflops": flops / seconds / 1e12,
                        "frac_tensor": flops / seconds / 1e12 / pk["tensor"]})
        if is_chamfer:
            lane_peak = 148 * 128 * pk.get("sm_max_mhz", 1965.0) * 1e6
            rec.update({"bound": "fp32-issue", "achieved": flops / seconds / 1e12, "peak": lane_peak / 1e12,
                        "unit": "T FP32 lane-ops/s", "frac": flops / seconds / lane_peak})
        elif flops is not None and flops / (pk["tensor"] * 1e12) > nbytes / (pk["hbm"] * 1e9):
            rec.update({"bound": "tensor", "achieved": rec["tflops"], "peak": pk["tensor"], "unit": "TFLOP/s", "frac": rec["frac_tensor"]})
        else:
            rec.update({"bound": "hbm", "achieved": rec["hbm_gbs"], "peak": pk["hbm"], "unit": "GB/s", "frac": rec["frac_hbm"]})
        if traffic is not None:
            rec["traffic"] = traffic
            rec["traffic_over_algorithmic"] = traffic / nbytes if nbytes else None
            if rec["bound"] == "hbm" and traffic < 0.5 * nbytes:
                rec["bound"] = "l2-latency"
        # a launch whose whole roofline time is below a microsecond cannot reach either roof: it is a chain of a few
        # dependent memory round trips plus launch latency (the cfg2-sized kernels); `large` holds roofline-sized shapes
        roof_us = max(nbytes / (pk["hbm"] * 1e9), (flops or 0.0) / (pk["tensor"] * 1e12) if not is_chamfer else 0.0) * 1e6
        rec["roofline_time_us"] = roof_us
        if roof_us < 1.0 and not is_chamfer:
            rec["bound"] = "latency"
        return rec
    return {"bound": "hbm", "achieved": None, "peak": pk["hbm"], "unit": "GB/s", "frac": None}


def ncu_traffic(workload: str, kernel_short: str):
    """DRAM bytes per launch of `kernel_short` from the committed ncu pass over this workload's step
    (profiles/traffic_<workload>.json, written by tools/ncu_traffic.py from `ncu --metrics
    dram__bytes_read.sum,dram__bytes_write.sum`); None when the file or the kernel is absent."""
    path = os.path.join(ROOT, "profiles", f"traffic_{workload}.json")
    if not os.path.exists(path):
        return None, None
    try:
        with open(path) as f:
            d = json.load(f)
    except Exception:
        return None, None
    base = kernel_short.replace("pcuda::", "").replace("tc::", "")
    for k, v in d.get("kernels", {}).items():
        if k == base or k.endswith(base) or base.endswith(k):
            return v.get("dram_bytes_per_launch"), f"profiles/traffic_{workload}.json ({d.get('source', 'ncu')})"
    return None, None


def kernel_profile(step, w: dict, pk: dict, reps: int = 5, record: bool = True, workload: str = ""):
    """Per-kernel device time of the (graph-replayed) step from CUPTI activity records, taken live in
    this process right after the timed region: which kernel dominates, and its roofline.
    Every rank must call this (the step holds the gradient all-reduce); only `record` ranks profile."""
    import torch
    if not record:
        for _ in range(reps):
            step.run()
        torch.cuda.synchronize()
        return None, None
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(reps):
            step.run()
        torch.cuda.synchronize()
    agg = {}
    for e in prof.events():
        if e.device_type != torch.autograd.DeviceType.CUDA:
            continue
        a = agg.setdefault(e.name, [0, 0.0])
        a[0] += 1
        a[1] += e.device_time
    if not agg:
        raise RuntimeError("no device activity records")
    tot = sum(v[1] for v in agg.values())
    ours = {k: v for k, v in agg.items() if "pcuda::" in k}
    top = sorted(ours.items(), key=lambda kv: -kv[1][1])
    short = lambda k: k.replace("(anonymous namespace)::", "").replace("void ", "").split("(")[0]
    shares = [{"kernel": short(k), "launches_per_step": v[0] / reps, "us_per_launch": v[1] / v[0],
               "share_of_step": v[1] / tot} for k, v in top[:8]]
    dom_name, dom = top[0]
    roof = {"kernel": short(dom_name), "launches_per_step": dom[0] / reps, "launch_us": dom[1] / dom[0],
            "share_of_step": dom[1] / tot, "device_busy_us_per_step": tot / reps,
            "libpcuda_share_of_step": sum(v[1] for v in ours.values()) / tot}
    traffic, tsrc = ncu_traffic(workload, short(dom_name))
    roof.update(kernel_roofline(dom_name, w, dom[1] / dom[0] * 1e-6, pk, traffic))
    roof["traffic"], roof["traffic_source"] = traffic, tsrc
    for sh, (k, v) in zip(shares, top[:8]):
        r = kernel_roofline(k, w, v[1] / v[0] * 1e-6, pk, ncu_traffic(workload, short(k))[0])
        sh.update({"bound": r.get("bound"), "frac": r.get("frac"), "frac_hbm": r.get("frac_hbm"), "frac_tensor": r.get("frac_tensor")})
    # the serial tail of the step (bucket fill, exchange, SGD): listed whatever their rank among the kernels
    roof["tail_kernels"] = [{"kernel": short(k), "launches_per_step": v[0] / reps, "us_per_launch": v[1] / v[0]}
                            for k, v in ours.items() if any(f in k for f in ("p2p_allreduce", "p2p_sum_f64", "sgd_momentum", "grad_sum_pack", "ncclDevKernel"))]
    roof["tail_kernels"] += [{"kernel": short(k), "launches_per_step": v[0] / reps, "us_per_launch": v[1] / v[0]}
                             for k, v in agg.items() if "nccl" in k.lower() and k not in ours]
    roof["peak_source"] = pk["source"]
    roof["note"] = ("dominant libpcuda kernel of the graph-replayed step at this workload (CUPTI device durations, warm L2, captured "
                    "without programmatic dependent launch so that a duration is the kernel's own work, not its wait); "
                    "at cfg2 sizes (2400 points per call) every kernel is latency-bound - `large` holds the roofline-sized shapes")
    return shares, roof


# ------------------------------------------------------------------------------------------------
def time_op(fn, iters: int, flush):
    import torch
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.fill_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return statistics.mean(ts)


def time_op_stream(fn, iters: int):
    """Mean device time per call of `iters` calls queued back to back (one CUDA-event interval): the launch and
    autograd overhead of a call overlaps the previous call's kernel, so this is the op's throughput.  Only for ops whose
    inputs are larger than the 126 MB L2 (nothing survives in cache from one call to the next; no flush needed)."""
    import torch
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def op_breakdown(step, w: dict, flush, pk: dict):
    """Per-op device time at this workload (cold L2) and the roofline of the dominant one."""
    import torch
    from pointcloududa_b200 import ops
    from pointcloududa_b200.utils.loss import batch_NN_loss, entropy_map
    I = step.inputs
    B, C, H, W, N = w["B"], w["C"], w["H"], w["W"], w["N"]
    elems = B * C * H * W
    act, norm, rp = w["activation"], w["normalize"], w["return_prob"]
    out = {}
    t = time_op(lambda: entropy_map(I["oS"], act, norm, return_prob=rp), 10, flush)
    out["entropy_fwd"] = dict(ms=t, bytes=(8 + (4 if rp else 0)) * elems)
    zr = I["oT"].detach().requires_grad_(True)
    mm = entropy_map(zr, act, norm)
    t = time_op(lambda: torch.autograd.grad(mm, zr, I["grad_mapT"], retain_graph=True), 10, flush)
    out["entropy_bwd"] = dict(ms=t, bytes=12 * elems)
    t = time_op(lambda: ops.chamfer_nn(I["vertS"], I["vertexA"]), 10, flush)
    out["chamfer_fwd"] = dict(ms=t, pairs=2 * B * N * N, bytes=2 * 24 * B * N)
    xr = I["vertS"].detach().requires_grad_(True)
    loss = batch_NN_loss(x=xr, y=I["vertexA"])
    t = time_op(lambda: torch.autograd.grad(loss, xr, retain_graph=True), 10, flush)
    out["chamfer_bwd"] = dict(ms=t, bytes=B * N * 2 * 36)
    d4 = step.d4
    pts = I["vertS"].transpose(2, 1)
    with torch.no_grad():
        t = time_op(lambda: d4.feat(pts), 10, flush)
    mac_pp = 278912 if not (w.get("feature_transform") or w.get("ext")) else None
    out["d4_feat_fwd"] = dict(ms=t, flops=2 * 278912 * B * N if mac_pp else None)
    for k, v in out.items():
        if v.get("bytes"):
            v["gbs"] = v["bytes"] / v["ms"] / 1e6
        if v.get("flops"):
            v["tflops"] = v["flops"] / v["ms"] / 1e9
        if v.get("pairs"):
            v["gpairs_s"] = v["pairs"] / v["ms"] / 1e6
    dom = max(out, key=lambda k: out[k]["ms"])
    d = out[dom]
    if dom.startswith("d4"):
        roof = {"bound": "tensor", "achieved": d["tflops"], "peak": pk["tensor"], "unit": "TFLOP/s",
                "frac": d["tflops"] / pk["tensor"], "traffic": None}
    else:
        roof = {"bound": "hbm", "achieved": d["gbs"], "peak": pk["hbm"], "unit": "GB/s",
                "frac": d["gbs"] / pk["hbm"], "traffic": None}
    roof["kernel"] = dom
    roof["launch_ms"] = d["ms"]
    roof["peak_source"] = pk["source"]
    roof["note"] = ("op-level duration (CUDA events, cold L2) of the dominant hot-path op at this workload; "
                    "cfg2-sized ops are launch-latency-bound, see `large` for the roofline-sized shapes")
    return out, roof


def large_shape_rooflines(flush, pk: dict):
    """Roofline-sized shapes (cfg-5 per-rank shard / cfg-4) where % of peak is meaningful."""
    import torch
    from pointcloududa_b200 import ops
    from pointcloududa_b200.synth import synth_clouds
    from pointcloududa_b200.utils.loss import entropy_map
    dev = flush.device
    res = {}
    # cfg-5 full batch (BASELINE configs[4]): 335 MB per tensor, well beyond the 126 MB L2
    z = torch.randn(256, 5, 256, 256, device=dev) * 3
    g = torch.randn_like(z)
    nb = z.numel() * 4
    # two timings per op: `single_shot` = one call between two events on an idle GPU with the L2 flushed (includes the
    # ~30 us of Python / autograd launch path in front of a ~150 us kernel); the headline `ms` = calls queued back to
    # back, inputs (335 MB each) larger than the L2, i.e. the op's sustained throughput
    t1 = time_op(lambda: entropy_map(z, "softmax", True), 10, flush)
    t = time_op_stream(lambda: entropy_map(z, "softmax", True), 20)
    res["entropy_fwd_B256_C5_256x256"] = dict(ms=t, gbs=2 * nb / t / 1e6, frac_hbm=2 * nb / t / 1e6 / pk["hbm"],
                                              single_shot=dict(ms=t1, frac_hbm=2 * nb / t1 / 1e6 / pk["hbm"]),
                                              l2="inputs 335 MB > L2 126 MB, calls back to back")
    zr = z.detach().requires_grad_(True)
    mm = entropy_map(zr, "softmax", True)
    t1 = time_op(lambda: torch.autograd.grad(mm, zr, g, retain_graph=True), 10, flush)
    t = time_op_stream(lambda: torch.autograd.grad(mm, zr, g, retain_graph=True), 20)
    res["entropy_bwd_B256_C5_256x256"] = dict(ms=t, gbs=3 * nb / t / 1e6, frac_hbm=3 * nb / t / 1e6 / pk["hbm"],
                                              single_shot=dict(ms=t1, frac_hbm=3 * nb / t1 / 1e6 / pk["hbm"]),
                                              l2="inputs 335 MB > L2 126 MB, calls back to back")
    del z, g, zr, mm
    torch.cuda.empty_cache()
    for B, N in ((64, 4096), (64, 16384)):
        x, y = synth_clouds(B, N, seed=5, mode="near")
        x, y = x.to(dev), y.to(dev)
        t = time_op(lambda: ops.chamfer_nn(x, y), 5, flush)
        # not HBM-bound by construction (24 B per point in + out): the binding resource is the FP32 pipe.
        # The prefiltered kernel issues 3 packed FFMA2 per two ordered pairs (3 lane-operations per pair) and
        # re-evaluates the reference's 5-operation formula only for pairs that can still lower a running
        # minimum; peak = SMs x 128 FP32 lanes x max SM clock.  `exact_formula_ceiling_tpairs_s` is what a
        # kernel that evaluates all 5 operations for every pair could reach at 100 % lane utilisation.
        pairs = 2.0 * B * N * N
        lane_peak = ops.sm_count() * 128 * pk.get("sm_max_mhz", 1965.0) * 1e6
        ent = dict(ms=t, tpairs_s=pairs / t / 1e9, hbm_gbs=48.0 * B * N / t / 1e6,
                   fp32_lane_ops_per_s=3.0 * pairs / (t * 1e-3),
                   frac_fp32_lanes=3.0 * pairs / (t * 1e-3) / lane_peak,
                   exact_formula_ceiling_tpairs_s=lane_peak / 5.0 / 1e12)
        # the exact packed kernel (every pair through the 5-operation formula) for comparison
        ops.tune(1, 200000 + (4128 if N >= 4096 else 0))
        try:
            t2 = time_op(lambda: ops.chamfer_nn(x, y), 3, flush)
        finally:
            ops.tune(1, 0)
        ent["exact_kernel"] = dict(ms=t2, tpairs_s=pairs / t2 / 1e9, frac_fp32_lanes=5.0 * pairs / (t2 * 1e-3) / lane_peak)
        res[f"chamfer_fwd_B{B}_N{N}"] = ent
    from pointcloududa_b200.networks.PointNetCls import PointNetCls
    net = PointNetCls().to(dev).train()
    pts = torch.rand(32, 4096, 3, device=dev).transpose(2, 1)
    with torch.no_grad():
        t = time_op(lambda: net.feat(pts), 5, flush)
    fl = 2 * 278912 * 32 * 4096
    res["d4_feat_fwd_B32_N4096"] = dict(ms=t, tflops=fl / t / 1e9, frac_tensor=fl / t / 1e9 / pk["tensor"],
                                        note="whole op, eager launches (CPU-launch-bound); kernel-level figures follow")
    # kernel-level: the tcgen05 GEMMs of that forward and of one backward, from CUPTI device durations
    try:
        from torch.profiler import ProfilerActivity, profile
        ptsg = torch.rand(32, 4096, 3, device=dev, requires_grad=True)
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(3):
                o = net(ptsg.transpose(2, 1))[0]
                o.sum().backward()
            torch.cuda.synchronize()
        agg = {}
        for e in prof.events():
            if e.device_type == torch.autograd.DeviceType.CUDA and "pcuda::" in e.name:
                a = agg.setdefault(e.name, [0, 0.0]); a[0] += 1; a[1] += e.device_time
        M = 32 * 4096
        for frag, name, flops in (("ws_kernel<1>", "mlp_pool_gemm_128x1024_fwd", 2.0 * M * 128 * 1024),
                                  ("ws_kernel<0>", "mlp_dense_gemm_64x128_fwd", 2.0 * M * 64 * 128),
                                  ("ws_kernel<3>", "mlp_pool_dgrad_lowrank", 2.0 * M * 128 * 128),
                                  ("ws_kernel<2>", "mlp_dgrad_128x64", 2.0 * M * 128 * 64)):
            for k, v in agg.items():
                if frag in k:
                    us = v[1] / v[0]
                    rr = kernel_roofline(k, dict(B=32, N=4096, C=5, H=256, W=256, return_prob=True), us * 1e-6, pk)
                    res[name + "_B32_N4096"] = dict(us=us, tflops=flops / us / 1e6, frac_tensor=flops / us / 1e6 / pk["tensor"],
                                                    frac_tensor_sustained=(flops / us / 1e6 / pk["tensor_sustained"]) if pk.get("tensor_sustained") else None,
                                                    bound=rr.get("bound"), hbm_gbs=rr.get("hbm_gbs"), frac_hbm=rr.get("frac_hbm"),
                                                    frac=rr.get("frac"), algorithmic_bytes=rr.get("algorithmic_bytes_per_launch"))
    except Exception as e:
        res["kernel_level_error"] = repr(e)
    return res


def measure_step(step, host_flat, result_host, flush, steps: int, world: int, dev):
    """(seconds of `steps` device-resident steps, seconds of `steps` end-to-end steps) on this rank.
    Device-resident: per-step CUDA events, L2 flushed before every step (outside the intervals).  End to end: the way
    a training loop feeds the step -- a copy stream uploads step i+1's inputs from pinned host memory into one of two
    staging sets while step i computes; the compute stream moves the staged set into the step's static input buffers
    (device-to-device), runs the step and reads the 32-byte result back -- timed as ONE interval over all steps
    (every upload, staging move, flush, compute and read-back inside it)."""
    import torch
    import torch.distributed as dist

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for e0, e1 in ev:
        flush.fill_(1.0)
        e0.record(); step.run(); e1.record()
    barrier()
    t_dev = sum(e0.elapsed_time(e1) for e0, e1 in ev) / 1e3

    copy_stream = torch.cuda.Stream(device=dev)
    staging = [torch.empty_like(step.inputs_flat) for _ in range(2)]
    staged = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    main = torch.cuda.current_stream(dev)

    def upload(i):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[i % 2])          # the staging set is free again
            staging[i % 2].copy_(host_flat, non_blocking=True)
            staged[i % 2].record(copy_stream)

    barrier()
    for ev_ in consumed:
        ev_.record(main)
    e2e0, e2e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush.fill_(1.0)
    e2e0.record()
    upload(0)
    for i in range(steps):
        if i + 1 < steps:
            upload(i + 1)
        main.wait_event(staged[i % 2])
        step.inputs_flat.copy_(staging[i % 2], non_blocking=True)
        consumed[i % 2].record(main)
        flush.fill_(1.0)
        r = step.run()
        result_host.copy_(r, non_blocking=True)
    e2e1.record()
    barrier()
    t_e2e = e2e0.elapsed_time(e2e1) / 1e3
    tt = torch.tensor([t_dev, t_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_dev, t_e2e = tt.tolist()
    return t_dev, t_e2e


def build_step(w: dict, rank: int, dev, precision: str, use_graph: bool, exchange: str = "auto", sync_bn: bool = False,
               share_trunk: bool = False):
    """AdversarialStep of workload `w` with its inputs loaded (per-rank synthetic shard).  share_trunk False: the three D4
    passes of the step each run their own forward, as the reference does (the configuration every headline number is
    measured in); True: the product's default, P4 reuses what P2 already computed on the same cloud."""
    import torch
    from pointcloududa_b200.step import RESULT_FIELDS, AdversarialStep, StepConfig
    cfg = StepConfig(B=w["B"], C=w["C"], H=w["H"], W=w["W"], N=w["N"], activation=w["activation"],
                     normalize=w["normalize"], return_prob=w["return_prob"], precision=precision, sync_bn=sync_bn)
    step = AdversarialStep(cfg, dev, seed=0, exchange=exchange, share_trunk=share_trunk)
    # one pinned host buffer holding a step's inputs in the step's flat layout: one H2D copy per step
    host_flat = torch.zeros(step.input_numel(), dtype=torch.float32).pin_memory()
    host = step.input_layout(host_flat)
    for k, v in make_host_inputs(w, rank, pinned=False).items():
        host[k].copy_(v)
    step.inputs_flat.copy_(host_flat)
    result_host = torch.empty(len(RESULT_FIELDS), dtype=torch.float32).pin_memory()
    return step, host_flat, result_host


def scale_lines(args, rank: int, world: int, dev, flush):
    """The multi-GPU configurations BASELINE.json names, measured in the same run as `value` so that the driver's
    1/2/4/8-GPU scaling runs carry them: cfg-5 (batch 256 over 8 GPUs = 32 clouds of 4096 points per rank; weak
    scaling, per-rank work fixed) and cfg-3 (MM-WHS, global batch 32 split over the ranks; strong scaling, 4 per
    rank on 8 GPUs), plus the cfg-4 Chamfer-only sweep (batch 64 split over the ranks)."""
    import torch
    out = {}
    steps = max(5, min(args.steps, 20))
    for name, scaling in (("cfg5_rank", "weak"), ("cfg3", "strong"), ("cfg3_syncbn", "strong")):
        sync_bn = name.endswith("_syncbn")          # cfg-3 with BatchNorm statistics over the clouds of all ranks
        if sync_bn and (world == 1 or args.exchange == "torch"):
            continue
        progress(f"scale line {name}")
        w = dict(WORKLOADS[name.replace("_syncbn", "")])
        if scaling == "strong":
            if w["B"] % world != 0 or w["B"] // world < 2:
                out[name] = {"skipped": f"global batch {w['B']} does not split into >= 2 clouds on each of {world} ranks"}
                continue
            w["B"] //= world
        try:
            step, host_flat, result_host = build_step(w, rank, dev, args.precision, not args.no_graph, args.exchange, sync_bn)
            if not args.no_graph:
                step.capture(warmup=2)
            for _ in range(3):
                step.run()
            torch.cuda.synchronize()
            t_dev, t_e2e = measure_step(step, host_flat, result_host, flush, steps, world, dev)
            units = world if scaling == "weak" else 1           # per-rank-shard steps (weak) / global-batch steps (strong)
            out[name] = {"scaling": scaling, "batchnorm": "cross-rank statistics" if sync_bn else "per-rank statistics", "per_rank_batch": w["B"], "global_batch": w["B"] * world, "steps": steps,
                         "workload": workload_label(name.replace("_syncbn", ""), w), "exchange": step.exchange,
                         "value": units * steps / t_dev,
                         "unit": "global-batch steps/s" if scaling == "strong" else "per-rank-shard steps/s summed over ranks",
                         "ms_per_step": t_dev / steps * 1e3, "samples_per_s": world * w["B"] * steps / t_dev,
                         "e2e": {"value": units * steps / t_e2e, "ms_per_step": t_e2e / steps * 1e3,
                                 "h2d_bytes_per_step": host_flat.numel() * 4, "d2h_bytes_per_step": result_host.numel() * 4}}
            step.close()
            del step, host_flat
            torch.cuda.empty_cache()
        except Exception as e:
            out[name] = {"error": repr(e)}
    try:
        out["cfg4_chamfer"] = chamfer_sweep(rank, world, dev, flush)
    except Exception as e:
        out["cfg4_chamfer"] = {"error": repr(e)}
    return out


CFG4_POINTS = (256, 512, 1024, 2048, 4096, 8192, 16384)


def chamfer_sweep(rank: int, world: int, dev, flush, B: int = 64, points=CFG4_POINTS, iters: int = 5):
    """BASELINE.json configs[3]: Chamfer-only sweep, 256-16k points per cloud, batch 64 sharded over the ranks
    (no collective on the data path: each rank owns 64 / n_gpus clouds).  Forward (both directions, indices, loss)
    per size; ordered pairs/s = 2 * B * N^2 / max-over-ranks device time, L2 flushed before every launch."""
    import torch
    import torch.distributed as dist
    from pointcloududa_b200 import ops
    from pointcloududa_b200.synth import synth_clouds
    if B % world:
        return {"skipped": f"batch {B} does not split over {world} ranks"}
    Br = B // world
    res = {"batch": B, "per_rank_batch": Br, "scaling": "strong", "unit": "T ordered pairs/s", "sizes": {}}
    for N in points:
        x, y = synth_clouds(B, N, seed=9, mode="near")
        x, y = x[rank * Br:(rank + 1) * Br].to(dev), y[rank * Br:(rank + 1) * Br].to(dev)
        for _ in range(2):
            ops.chamfer_nn(x, y)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ts = []
        for _ in range(iters):
            flush.fill_(1.0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ops.chamfer_nn(x, y); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = torch.tensor([statistics.mean(ts)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
        res["sizes"][str(N)] = {"ms": ms, "tpairs_s": 2.0 * B * N * N / ms / 1e9}
    return res


def run_ours(args):
    import torch
    import torch.distributed as dist
    from pointcloududa_b200 import dist as pdist
    from pointcloududa_b200 import ops
    from pointcloududa_b200.step import RESULT_FIELDS

    rank, local, world = pdist.init_from_env("nccl")
    if world != args.gpus:
        if rank == 0:
            print(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}; using WORLD_SIZE", file=sys.stderr)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = pdist.bind_to_gpu_numa(local) if world > 1 else None      # before any pinned host allocation
    w = dict(WORKLOADS[args.workload])
    pk = peaks()
    step, host_flat, result_host = build_step(w, rank, dev, args.precision, not args.no_graph, args.exchange)
    h2d_bytes = host_flat.numel() * 4
    flush = torch.empty(FLUSH_BYTES // 4, device=dev)

    # launches per step, counted on an eager step (a graph replays exactly these)
    step.run(); torch.cuda.synchronize()
    l0 = ops.launch_count(); step.run(); torch.cuda.synchronize()
    launches_per_step = ops.launch_count() - l0

    use_graph = not args.no_graph
    if use_graph:
        step.capture(warmup=2)
    for _ in range(max(args.warmup, 3)):
        step.run()
    torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    progress("timed region")
    t_dev, t_e2e = measure_step(step, host_flat, result_host, flush, args.steps, world, dev)
    clocks = sampler.stop() if rank == 0 else None
    progress("kernel profile")
    last = {k: float(v) for k, v in zip(RESULT_FIELDS, result_host.tolist())}

    # kernel-level profile of the step: collective inside, so every rank runs the steps; rank 0 records
    kernel_shares, roof = None, None
    try:
        if use_graph:
            # With programmatic dependent launch a kernel is resident (and its CUPTI record runs) while it still waits
            # for its predecessor: the per-kernel durations of the profile are taken from a capture without it.
            ops.tune(8, 1)
            step.capture(warmup=1)
        kernel_shares, roof = kernel_profile(step, w, pk, record=(rank == 0), workload=args.workload)
    except Exception as e:       # CUPTI unavailable: keep the op-level roofline
        kernel_shares = {"error": repr(e)}
        if rank == 0 and world > 1:
            for _ in range(5):   # keep the collective count matched with the other ranks
                step.run()
    finally:
        ops.tune(8, 0)
    # the multi-GPU configurations BASELINE.json names + the Chamfer-only sweep (collective: every rank takes part)
    scale = None
    if not args.skip_scale:
        progress("scale lines")
        scale = scale_lines(args, rank, world, dev, flush)
    progress("shared-trunk variant")
    # the same step with the product's default forward sharing (collective inside: every rank runs it)
    shared = None
    try:
        step2, hf2, rh2 = build_step(w, rank, dev, args.precision, use_graph, args.exchange, share_trunk=True)
        step2.run(); torch.cuda.synchronize()
        l0 = ops.launch_count(); step2.run(); torch.cuda.synchronize()
        launches2 = ops.launch_count() - l0
        if use_graph:
            step2.capture(warmup=2)
        for _ in range(3):
            step2.run()
        torch.cuda.synchronize()
        k2 = max(5, min(args.steps, 30))
        t2_dev, t2_e2e = measure_step(step2, hf2, rh2, flush, k2, world, dev)
        shared = {"value": world * k2 / t2_dev, "unit": UNIT, "ms_per_step": t2_dev / k2 * 1e3, "steps": k2,
                  "e2e": {"value": world * k2 / t2_e2e, "ms_per_step": t2_e2e / k2 * 1e3},
                  "gpu_launches_per_step": int(launches2),
                  "what": "P4 takes both trunks and the STN head of D4's forward pass from P2 (same cloud, same weights: "
                          "networks/PointNetCls.py share_forward); tests/test_gpu_step_parity.py checks bit-identity"}
        if world > 1:
            step2.close()
        del step2
    except Exception as e:
        shared = {"error": repr(e)}
    line = None
    progress("rank-0 extras (op breakdown, large shapes, baselines)")
    if rank == 0:
        breakdown, roof_ops = op_breakdown(step, w, flush, pk)
        if roof is None:
            roof = roof_ops
        else:
            roof["op_level"] = roof_ops
        large = None
        if not args.skip_large:
            try:
                large = large_shape_rooflines(flush, pk)
            except Exception as e:  # never lose the headline line to an auxiliary measurement
                large = {"error": repr(e)}
        cpu = None
        if world == 1 and not args.skip_cpu:
            try:
                cpu = run_cpu_baseline(w, budget_s=args.cpu_budget)
            except Exception as e:
                cpu = {"error": repr(e)}
        eager = None
        if not args.skip_eager:
            try:
                eager = run_gpu_eager_baseline(w, dev, make_host_inputs(w, rank, pinned=False), flush,
                                               steps=min(max(args.steps, 5), 30))
                if "value" in eager:
                    eager["ours_over_eager_value"] = (args.steps / t_dev) / eager["value"]     # per GPU, inputs resident
            except Exception as e:
                eager = {"error": repr(e)}
        line = {
            "metric": METRIC, "value": world * args.steps / t_dev, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": t_dev / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32" if args.precision == "fp32" else "f32+bf16",
            "data": "synthetic",
            "config": bench_config(args.workload, w, world, args.precision, use_graph),
            "e2e": {"value": world * args.steps / t_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": result_host.numel() * 4, "ms_per_step": t_e2e / args.steps * 1e3},
            "gpu_launches": int(launches_per_step * args.steps),
            "gpu_launches_per_step": int(launches_per_step),
            "samples_per_s": world * w["B"] * args.steps / t_dev,
            "exchange": step.exchange, "host_affinity": numa, "scale": scale, "shared_trunk": shared,
            "parity": {"chamfer_indices": "bit-exact (int64 equality vs golden / oracle / torch on this GPU)",
                       "fp32": "entropy 1e-5 rel (+5e-7 abs), Chamfer loss 2e-6, D4 fp32 path 2e-5 values / 2e-4 gradients vs float64",
                       "bf16_mlp": "op level (tensor-core shared-MLP stack, every gradient, L2 and max-norm) <= 1e-2 vs the float64 "
                                   "same-rounding oracle; whole network / whole step: max(1e-2, 4x the float32-vs-float64 distance of "
                                   "the oracle's own arithmetic), measured in the test (DESIGN.md §2)",
                       "tests": "tests/test_gpu_parity.py, tests/test_gpu_step_parity.py, tests/test_gpu_multi.py"},
            "roofline": roof, "kernels": kernel_shares, "ops": breakdown, "large": large, "cpu_baseline": cpu, "gpu_eager_baseline": eager, "clocks": clocks,
            "last_step": last,
        }
        print(json.dumps(line))
    progress("done")
    if world > 1:
        dist.barrier()
        step.close()
        dist.destroy_process_group()
    return line


def run_reference(args):
    """Reference arm for this tier: the reference's own CPU path — its unmodified PyTorch code from baseline/_ref
    under torch eager on every host core (oracle port only if baseline/_ref is not installed) — on the GPU arm's
    workload / metric / unit, `--steps` timed steps after `--warmup` untimed ones.  Rank 0 alone runs."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = dict(WORKLOADS[args.workload])
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    warm = max(args.warmup, 1)
    cpu = run_cpu_baseline(w, budget_s=0.0, min_steps=args.steps, max_steps=args.steps, warmup=warm)
    v = 1e3 / cpu["ms_per_step_mean"]            # K steps / total time of the K steps
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": warm, "ms_per_step": 1e3 / v, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": bench_config(args.workload, w, world, args.precision, not args.no_graph),
        "cpu_baseline": {**cpu, "value": v},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def run_chamfer_only(args):
    """`--workload cfg4` (BASELINE.json configs[3]): the Chamfer-only sweep as its own line — metric = ordered pairs/s
    at the largest size, every size under `sweep`; batch 64 sharded over the ranks, no data-path collective."""
    import torch
    import torch.distributed as dist
    from pointcloududa_b200 import dist as pdist
    from pointcloududa_b200 import ops
    rank, local, world = pdist.init_from_env("nccl")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    flush = torch.empty(FLUSH_BYTES // 4, device=dev)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    n0 = ops.launch_count()
    sweep = chamfer_sweep(rank, world, dev, flush, iters=max(3, min(args.steps, 20)))
    launches = ops.launch_count() - n0
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0 and "sizes" in sweep:
        top = sweep["sizes"][str(CFG4_POINTS[-1])]
        pk = peaks()
        lane_peak = 148 * 128 * pk.get("sm_max_mhz", 1965.0) * 1e6
        line = {"metric": "chamfer_ordered_pairs_per_s", "value": top["tpairs_s"] * 1e12, "unit": "ordered pairs/s", "n_gpus": world,
                "steps": max(3, min(args.steps, 20)), "warmup": 2, "ms_per_step": top["ms"], "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"cfg4: Chamfer-only forward (both directions, indices, loss), batch 64 over {world} rank(s), "
                                       f"{CFG4_POINTS[-1]} points per cloud (sweep 256-16384 under `sweep`)",
                           "l2": "flushed before every launch"},
                "sweep": sweep, "gpu_launches": int(launches), "clocks": clocks,
                "roofline": {"bound": "fp32-issue", "achieved": 3.0 * top["tpairs_s"], "peak": lane_peak * world / 1e12,
                             "unit": "T FP32 lane-ops/s (3 per ordered pair: the prefilter)", "frac": 3.0 * top["tpairs_s"] * 1e12 / (lane_peak * world),
                             "traffic": None}}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    # PCUDA_BENCH_WATCHDOG=<seconds>: dump every thread's Python stack and exit if the run is still going by then
    wd = float(os.environ.get("PCUDA_BENCH_WATCHDOG", "0") or 0)
    if wd > 0:
        import faulthandler
        faulthandler.dump_traceback_later(wd, exit=True)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS) + ["cfg4"])
    ap.add_argument("--precision", default="bf16", choices=["fp32", "bf16"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-large", action="store_true")
    ap.add_argument("--skip-eager", action="store_true")
    ap.add_argument("--skip-scale", action="store_true")
    ap.add_argument("--exchange", default="auto", choices=["auto", "p2p", "nccl", "torch"],
                    help="how D4's gradient bucket is summed over the ranks (pointcloududa_b200/step.py)")
    ap.add_argument("--cpu-budget", type=float, default=15.0)
    args = ap.parse_args()
    if args.impl == "reference":
        if args.workload == "cfg4":
            args.workload = "cfg2"
        run_reference(args)
    elif args.workload == "cfg4":
        run_chamfer_only(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
