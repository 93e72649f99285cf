"""Warp-stall samples per CUDA source line from `ncu --page source --csv --print-source cuda,sass` (top N lines)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = None
agg = {}
fname = ""
for r in rows:
    if r and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    if r and r[0] == "Line No":
        hdr = r
        si = hdr.index("# Samples")
        continue
    if hdr is None or len(r) <= si or not r[0].strip().isdigit():
        continue
    if r[2] != "-":  # SASS row under a CUDA line
        continue
    key = (fname, int(r[0]), r[1].strip()[:90])
    agg[key] = agg.get(key, 0) + (int(r[si]) if r[si].strip().isdigit() else 0)
tot = sum(agg.values())
print("total samples", tot)
for (f, ln, src), v in sorted(agg.items(), key=lambda kv: -kv[1])[:n]:
    print(f"{v:7d} {100*v/tot:5.1f}%  {f}:{ln}  {src}")
