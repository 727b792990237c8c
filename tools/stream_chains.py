"""Per-stream view of the graph-replayed step from CUPTI records: for every kernel its 'effective cost' = its end minus the
end of the kernel before it on the same stream (what it adds to that stream's chain; with programmatic dependent launch a
kernel's own duration includes the time it spent waiting for its predecessor).  Aggregated by kernel name per stream.
python tools/stream_chains.py [cfg2]"""
import os, sys, collections
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from torch.profiler import profile, ProfilerActivity

wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
dev = torch.device("cuda", 0)
step, hf, rh = bench.build_step(dict(bench.WORKLOADS[wl]), 0, dev, "bf16", True)
step.capture(warmup=2)
flush = torch.empty(bench.FLUSH_BYTES // 4, device=dev)
for _ in range(5):
    step.run()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    flush.fill_(1.0)
    torch.cuda.synchronize()
    step.run()
    torch.cuda.synchronize()
short = lambda n: n.replace("(anonymous namespace)::", "").replace("void ", "").replace("pcuda::", "").split("(")[0][:52]
evs = []
for e in prof.profiler.kineto_results.events():
    if str(e.device_type()).endswith("CUDA") and e.duration_ns() > 0 and "fill" not in e.name().lower()[:0]:
        evs.append((e.start_ns() / 1e3, (e.start_ns() + e.duration_ns()) / 1e3, e.device_resource_id(), short(e.name())))
evs.sort()
evs = [e for e in evs if e[0] >= evs[1][0]] if len(evs) > 2 else evs       # drop the flush fill
t0 = min(e[0] for e in evs)
print(f"{wl}: {len(evs)} device records, span {max(e[1] for e in evs) - t0:.1f} us (under the profiler)")
streams = collections.defaultdict(list)
for e in evs:
    streams[e[2]].append(e)
for sid, lst in sorted(streams.items(), key=lambda kv: -len(kv[1])):
    lst.sort(key=lambda e: e[1])
    agg = collections.defaultdict(lambda: [0, 0.0])
    prev_end = lst[0][0]
    total = 0.0
    for (s, en, _, name) in lst:
        eff = en - max(prev_end, min(s, prev_end))       # end-to-end delta
        eff = en - prev_end if en > prev_end else 0.0
        agg[name][0] += 1; agg[name][1] += eff
        total += eff
        prev_end = max(prev_end, en)
    print(f"-- stream {sid}: {len(lst)} kernels, first start +{lst[0][0] - t0:.1f} us, last end +{max(e[1] for e in lst) - t0:.1f} us, sum of effective costs {total:.1f} us")
    for name, (n, c) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
        print(f"   {c:7.1f} us {n:3d} x {c / n:6.2f}  {name}")
