"""SASS instruction histogram of the tcgen05 / TMA kernels in libicf_b200.so (run where cuobjdump is installed).
usage: python tools/sass_histogram.py [path/to/libicf_b200.so] > profiles/rNN_sass_histogram.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "imagecfgen-pytorch_b200", "icf_b200", "libicf_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
names = {}
cur = None
hist = collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        hist[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        hist[cur][m.group(1)] += 1
        hist[cur]["_total"] += 1
dem = subprocess.run(["c++filt"], input="\n".join(hist), capture_output=True, text=True).stdout.splitlines()
cols = ["UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "REDG", "ATOMG"]
print("# SASS instruction histogram of the tcgen05 / TMA kernels in libicf_b200.so (cuobjdump -sass, sm_100a), per kernel instance:")
print("# UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG / UTMASTG = TMA tensor load / store, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops,")
print("# REDG / ATOMG = global reductions / atomics")
print(f"{'kernel':64s} {'instr':>6s} " + " ".join(f"{c:>8s}" for c in cols))
for (mangled, h), d in zip(hist.items(), dem):
    if not (h["UTCHMMA"] or h["UTMALDG"]):
        continue
    name = re.sub(r"\(anonymous namespace\)::", "", d)
    name = re.sub(r"^void ", "", name).split("(")[0]
    print(f"{name[:64]:64s} {h['_total']:6d} " + " ".join(f"{h[c]:8d}" for c in cols))
