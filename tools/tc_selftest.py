"""Tensor-core shared-MLP self-test: every tcgen05 piece alone and all together against the FP32
CUDA-core kernels of the same library, on several stacks / shapes.

    python tools/tc_selftest.py            # driver: one subprocess per (case, mask) with a timeout
    python tools/tc_selftest.py CASE MASK  # worker

MASK is the TUNE_MLP_TC_MASK value: set bits switch pieces BACK to FP32
(1 forward, 2 pooled dgrad, 4 dense dgrad, 8 wgrad, 16 Gram); 0 = everything on tensor cores.
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = [
    # name, chans, relus, B, N, pool
    ("stn_b4_n64", [3, 64, 128, 1024], [1, 1, 1], 4, 64, True),
    ("trunk_b8_n300", [3, 64, 128, 1024], [1, 1, 0], 8, 300, True),
    ("trunk_b2_n1000", [3, 64, 128, 1024], [1, 1, 0], 2, 1000, True),
    ("stnkd_b3_n70", [64, 64, 128, 1024], [1, 1, 1], 3, 70, True),
    ("ftpost_b4_n100", [64, 128, 1024], [1, 0], 4, 100, True),
    ("ext_b3_n40", [3, 8, 64, 128, 256, 512, 1024], [1, 1, 1, 1, 0, 1], 3, 40, True),
    ("trunk_b32_n4096", [3, 64, 128, 1024], [1, 1, 0], 32, 4096, True),
]
MASKS = [30, 29, 27, 23, 15, 0]


def worker(case_name: str, mask: int) -> None:
    import numpy as np
    import torch
    import torch.nn as nn
    from pointcloududa_b200 import ops
    from pointcloududa_b200.networks.PointNetCls import shared_mlp

    name, chans, relus, B, N, pool = next(c for c in CASES if c[0] == case_name)
    dev = "cuda:0"
    rs = np.random.RandomState(abs(hash(name)) % (2 ** 31))
    torch.manual_seed(1)
    convs, bns = [], []
    for l in range(len(chans) - 1):
        conv = nn.Conv1d(chans[l], chans[l + 1], 1).to(dev)
        bn = nn.BatchNorm1d(chans[l + 1]).to(dev)
        with torch.no_grad():
            bn.weight.copy_(torch.from_numpy(1.0 + 0.5 * rs.standard_normal(chans[l + 1])).float())
            bn.bias.copy_(torch.from_numpy(0.2 * rs.standard_normal(chans[l + 1])).float())
        convs.append(conv); bns.append(bn)
    pts = torch.from_numpy(rs.rand(B, N, chans[0])).float().to(dev)
    res = {}
    outs = {}
    # three evaluations: FP32 kernels; tensor-core forward + FP32 backward (mask 30); the mask under test.
    # Forward quantities are compared with FP32.  Backward quantities are compared with the mask-30 run:
    # both back-propagate through the SAME tensor-core forward (same arg-max selection, same saved
    # statistics), so they must agree to bf16 rounding, whereas a bf16 forward may legitimately select
    # a different, nearly tied arg-max point than the FP32 forward.
    for tag, prec, mk in (("fp32", "fp32", 0), ("tcfwd", "bf16", 30), ("test", "bf16", mask)):
        ops.tune(3, mk)
        for m in convs + bns:
            for p in m.parameters():
                p.grad = None
        for bn in bns:
            bn.reset_running_stats()
        x = pts.transpose(2, 1).detach().requires_grad_(True)
        out = shared_mlp(x, convs, bns, [bool(r) for r in relus], pool=pool, precision=prec)
        w = torch.from_numpy(np.random.RandomState(7).standard_normal(tuple(out.shape))).float().to(dev)
        (out * w).sum().backward()
        torch.cuda.synchronize()
        g = {"out": out.detach().clone(), "gx": x.grad.clone()}
        for l, (c, b) in enumerate(zip(convs, bns)):
            g[f"gW{l}"] = c.weight.grad.clone()
            g[f"ggamma{l}"] = b.weight.grad.clone()
            g[f"gbeta{l}"] = b.bias.grad.clone()
            g[f"rm{l}"] = b.running_mean.clone()
            g[f"rv{l}"] = b.running_var.clone()
        outs[tag] = g
    ops.tune(3, 0)
    for k in outs["fp32"]:
        fwd_q = k == "out" or k.startswith(("rm", "rv"))
        ref = outs["fp32"] if fwd_q else outs["tcfwd"]
        a, b = outs["test"][k].double(), ref[k].double()
        denom = b.abs().max().item() + 1e-30
        nrm = b.norm().item()
        if k.startswith(("ggamma", "gbeta")):
            # can be mathematically zero (BN without ReLU feeding a train-mode BN): use the scale of
            # the same layer's weight gradient
            wk = ref["gW" + k.lstrip("gamebt")].double()
            denom = max(denom, 1e-2 * wk.abs().max().item())
            nrm = max(nrm, 1e-2 * wk.norm().item() / max(1.0, (wk.numel() / b.numel()) ** 0.5))
        res[k] = dict(max=float((a - b).abs().max().item() / denom),
                      l2=float(((a - b).norm() / (nrm + 1e-30)).item()),
                      finite=bool(torch.isfinite(a).all().item()))
    print("RESULT " + json.dumps({"case": name, "mask": mask, "err": res}))


def driver() -> None:
    only = sys.argv[2:] if len(sys.argv) > 2 and sys.argv[1] == "--cases" else None
    masks = [int(m) for m in os.environ.get("TC_MASKS", "").split(",") if m] or MASKS
    bad = 0
    for case in CASES:
        if only and case[0] not in only:
            continue
        for mask in masks:
            try:
                r = subprocess.run([sys.executable, __file__, case[0], str(mask)], capture_output=True, text=True, timeout=180)
            except subprocess.TimeoutExpired:
                print(f"{case[0]:18s} mask={mask:2d} TIMEOUT")
                bad += 1
                continue
            line = next((ln for ln in r.stdout.splitlines() if ln.startswith("RESULT ")), None)
            if line is None:
                print(f"{case[0]:18s} mask={mask:2d} FAILED rc={r.returncode}: {(r.stderr or r.stdout).strip().splitlines()[-1:]}")
                bad += 1
                continue
            err = json.loads(line[7:])["err"]
            worst_l2 = max(v["l2"] for v in err.values())
            worst_max = max(v["max"] for v in err.values())
            wk = max(err, key=lambda k: err[k]["l2"])
            fin = all(v["finite"] for v in err.values())
            flag = "ok " if (fin and worst_l2 < 3e-2) else "BAD"
            if flag == "BAD":
                bad += 1
            print(f"{case[0]:18s} mask={mask:2d} {flag} worst l2={worst_l2:.2e} ({wk}) max={worst_max:.2e}  "
                  f"out={err['out']['l2']:.1e} gx={err['gx']['l2']:.1e} gW_last={err['gW%d' % (len(case[1]) - 2)]['l2']:.1e}")
    print(f"tc_selftest: {bad} bad")


if __name__ == "__main__":
    if len(sys.argv) == 3 and sys.argv[1] != "--cases":
        worker(sys.argv[1], int(sys.argv[2]))
    else:
        driver()
