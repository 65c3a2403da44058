"""Aggregate an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list by kernel."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[hi]
idx = {h: i for i, h in enumerate(hdr)}
per = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    name = r[idx["Kernel Name"]]
    name = name[name.find("::") + 2:] if "<unnamed>::" in name else name
    name = name.split("(")[0][:60]
    e = per.setdefault(r[idx["ID"]], {"name": name, "t": 0.0, "b": 0.0})
    v, u, m = float(r[idx["Metric Value"]].replace(",", "")), r[idx["Metric Unit"]], r[idx["Metric Name"]]
    if m.startswith("gpu__time"):
        e["t"] += v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(u, 1e-3)
    else:
        e["b"] += v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
agg = collections.OrderedDict()
for e in per.values():
    a = agg.setdefault(e["name"], [0, 0.0, 0.0])
    a[0] += 1
    a[1] += e["t"]
    a[2] += e["b"]
print(f"# {sys.argv[1]}: ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none over "
      f"{len(per)} launches (~3 steps, eager, cold cache)")
print(f"{'kernel':60s} {'launches':>8s} {'total us':>10s} {'DRAM MB/launch':>15s} {'GB/s':>8s}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:36]:
    print(f"{k:60s} {v[0]:8d} {v[1]:10.1f} {v[2] / v[0] / 1e6:15.2f} {v[2] / max(v[1], 1e-9) / 1e3:8.1f}")
