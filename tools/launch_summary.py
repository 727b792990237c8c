"""Summarise an ncu launch list (gpu__time_duration.sum CSV): one step of bench.py --no-graph, per kernel."""
import collections, csv, re, sys
f = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = [r for r in csv.reader(open(f)) if len(r) > 5]
hdr = rows[0]; ik = hdr.index("Kernel Name"); iv = hdr.index("Metric Value"); iid = hdr.index("ID")
im = hdr.index("Metric Name")
# (a list taken together with other metrics, e.g. the DRAM-traffic pass, holds several rows per launch)
data = [(int(r[iid]), r[ik], float(r[iv].replace(",", ""))) for r in rows[1:] if r[iid].isdigit() and r[im] == "gpu__time_duration.sum"]
def short(k):
    k = re.sub(r"\(.*", "", k)
    return k.replace("void ", "").replace("pcuda::", "").replace("<unnamed>::", "").replace("unnamed>::", "")[:100]
ch = [i for i, (a, k, v) in enumerate(data) if "chamfer_nn" in k]
# a step holds two Chamfer forwards (P1 source, P2 target): take the span between the 3rd and 5th
s, e = ch[2], ch[4]
step = data[s:e]
agg = collections.defaultdict(lambda: [0, 0.0])
for i, k, v in step:
    agg[short(k)][0] += 1; agg[short(k)][1] += v
tot = sum(v[1] for v in agg.values())
lib = lambda k: not k.startswith(("at::", "cutlass", "cublas", "gemm", "epilogue", "std::", "gemv", "magma", "internal"))
ours = sum(v[1] for k, v in agg.items() if lib(k)); n_ours = sum(v[0] for k, v in agg.items() if lib(k))
print(f"{f}: one step = {len(step)} launches, {tot/1e3:.1f} us summed (ncu: serialised, cold); "
      f"libpcuda {n_ours} launches {ours/1e3:.1f} us, framework {len(step)-n_ours} launches {(tot-ours)/1e3:.1f} us")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"  {v[1]/1e3:9.1f} us {v[0]:4d} x {v[1]/v[0]/1e3:7.1f} us {100*v[1]/tot:5.1f}%  {k}")
