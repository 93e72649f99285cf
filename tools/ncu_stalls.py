"""Totals of warp-stall reasons over an `ncu --page source --csv --print-source sass` dump (optionally a line range)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
si = hdr.index("# Samples")
data = [r for r in rows[2:] if len(r) > si and r[si].strip().isdigit()]
a = int(sys.argv[2]) if len(sys.argv) > 2 else 0
b = int(sys.argv[3]) if len(sys.argv) > 3 else len(data)
cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[si]) for r in data[a:b])
out = []
for c in cols:
    out.append((sum(int(r[c]) if r[c].strip().isdigit() else 0 for r in data[a:b]), hdr[c]))
print("samples", tot)
for v, k in sorted(out, reverse=True)[:12]:
    print(f"  {k:28s} {v:7d} {100 * v / max(tot, 1):5.1f}%")
