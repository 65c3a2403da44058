"""Top stalled SASS instructions of an `ncu --page source --csv` dump (development helper)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[0] != "Address"]
tot = sum(int(r[idx["# Samples"]]) for r in data)
print("total samples", tot, "instructions", len(data))
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
order = sorted(range(len(data)), key=lambda i: -int(data[i][idx["# Samples"]]))[:n]
for i in order:
    r = data[i]
    st = sorted(((int(r[idx[c]]), c[6:]) for c in stall_cols), reverse=True)[:2]
    print(f"{i:5d} {r[idx['# Samples']]:>7s} {r[idx['Instructions Executed']]:>9s}  {r[idx['Source']].strip()[:72]:72s} {st}")
