"""Condense an ncu CSV (`--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv`)
of one bench step into profiles/traffic_<workload>.json: per kernel, launches, mean duration and mean DRAM
bytes per launch.  bench.py reads that file to fill `roofline.traffic` for the dominant kernel.

    python tools/ncu_traffic.py gpurun_out/traffic_cfg2.csv cfg2 profiles/traffic_cfg2.json
"""
import csv
import json
import re
import sys


def short(name: str) -> str:
    name = name.replace("void ", "").split("(")[0]
    name = re.sub(r"pcuda::|tc::|<unnamed>::|\(anonymous namespace\)::|unnamed>::", "", name)
    name = re.sub(r"\(.*?Mode\)", "", name)
    return name.strip()


def main(src: str, workload: str, dst: str) -> None:
    rows = []
    with open(src, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    col = {h: i for i, h in enumerate(hdr)}
    agg = {}
    for r in rd:
        if len(r) != len(hdr):
            continue
        k = short(r[col["Kernel Name"]])
        metric, unit, val = r[col["Metric Name"]], r[col["Metric Unit"]], float(r[col["Metric Value"]].replace(",", ""))
        a = agg.setdefault(k, {"ids": set(), "ns": 0.0, "rd": 0.0, "wr": 0.0})
        a["ids"].add(r[col["ID"]])
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1.0)
        if metric == "gpu__time_duration.sum":
            a["ns"] += val * scale
        elif metric == "dram__bytes_read.sum":
            a["rd"] += val * scale
        elif metric == "dram__bytes_write.sum":
            a["wr"] += val * scale
    out = {"workload": workload, "source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum, "
                                           "--clock-control none, cold caches, serialised; file " + src,
           "kernels": {}}
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ns"]):
        n = len(a["ids"])
        out["kernels"][k] = {"launches": n, "us_per_launch": a["ns"] / n / 1e3,
                             "dram_read_bytes_per_launch": a["rd"] / n, "dram_write_bytes_per_launch": a["wr"] / n,
                             "dram_bytes_per_launch": (a["rd"] + a["wr"]) / n}
    with open(dst, "w") as f:
        json.dump(out, f, indent=1)
    print(f"{dst}: {len(out['kernels'])} kernels")


if __name__ == "__main__":
    main(*sys.argv[1:4])
