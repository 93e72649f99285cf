"""Prints the hottest SASS lines (warp-stall samples) of an `ncu --page source --csv --print-source sass` dump."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
si, src, ex = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = [r for r in rows[2:] if len(r) > si and r[si].strip().isdigit()]
tot = sum(int(r[si]) for r in data)
print("total samples", tot, " instructions", len(data))
top = sorted(range(len(data)), key=lambda i: -int(data[i][si]))[: int(sys.argv[2]) if len(sys.argv) > 2 else 30]
for i in sorted(top):
    r = data[i]
    st = sorted(((int(r[c]) if r[c].strip().isdigit() else 0, hdr[c]) for c in stall_cols), reverse=True)[:2]
    print(f"{i:5d} {int(r[si]):7d} {int(r[si]) / tot * 100:5.1f}%  exec={r[ex]:>9s}  {r[src][:70]:70s} {st[0][1]}:{st[0][0]} {st[1][1]}:{st[1][0]}")
