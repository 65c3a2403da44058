"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (for profiles/)."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[hi]
idx = {h: i for i, h in enumerate(hdr)}
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    name = r[idx["Kernel Name"]]
    name = name[name.find("::") + 2:] if "<unnamed>::" in name else name
    name = name.split("(")[0][:70]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += float(r[idx["Metric Value"]]) / 1e3
tot = sum(v[1] for v in agg.values())
n = sum(v[0] for v in agg.values())
print(f"# {sys.argv[1]}: {n} launches, {tot / 1e3:.2f} ms of kernel time (cold-cache, serialised: compare SHARES)")
print(f"{'kernel':70s} {'launches':>8s} {'total us':>10s} {'share':>7s} {'avg us':>8s}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{k:70s} {v[0]:8d} {v[1]:10.1f} {100 * v[1] / tot:6.1f}% {v[1] / v[0]:8.1f}")
