"""profiles/r2_sass_counts.txt: per-kernel counts of the SASS mnemonics that prove what each kernel is built on
(tcgen05 MMA = UTCHMMA, TMEM loads = LDTM, bulk copies = UBLKCP, tensor-core barriers = UTCBAR, packed FP32 =
FFMA2 / FADD2 / FMUL2, 3-input min/max = FMNMX3) from `cuobjdump -sass` of the built libpcuda.so."""
import collections
import re
import subprocess
import sys

LIB = sys.argv[1] if len(sys.argv) > 1 else "pointcloududa_b200/csrc/libpcuda.so"
MN = ("UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UBLKCP", "UTCBAR", "UTMALDG", "FFMA2", "FADD2", "FMUL2", "FMNMX3", "HMMA", "MUFU", "RED", "ATOM",
      "SYNCS", "DADD", "DFMA", "DMUL")
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
cur, table = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = name.replace("(anonymous namespace)::", "").replace("void ", "")
        name = re.sub(r"\(.*", "", name).replace("pcuda::", "")
        cur = table.setdefault(name, collections.Counter())
        continue
    if cur is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m:
        op = m.group(1)
        cur["total"] += 1
        for k in MN:
            if op.startswith(k):
                cur[k] += 1
print(f"# cuobjdump -sass {LIB}: instructions per kernel (static counts)")
print("kernel".ljust(58) + "total".rjust(7) + "".join(k.rjust(8) for k in MN))
tot = collections.Counter()
for name, c in table.items():
    tot.update(c)
    print(name[:57].ljust(58) + str(c["total"]).rjust(7) + "".join((str(c[k]) if c[k] else ".").rjust(8) for k in MN))
print("ALL".ljust(58) + str(tot["total"]).rjust(7) + "".join(str(tot[k]).rjust(8) for k in MN))
