"""Top SASS instructions by warp-stall samples from `ncu --page source --csv --print-source cuda,sass`:
address, samples, dominant stall reasons, the CUDA line it belongs to and the SASS text."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 50
hdr = None
items = []
cur_line, cur_src, fname = None, "", ""
for r in rows:
    if r and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    if r and r[0] == "Line No":
        hdr = r
        si = hdr.index("# Samples")
        stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if hdr is None or len(r) <= si:
        continue
    if r[0].strip().isdigit() and r[2] == "-":
        cur_line, cur_src = int(r[0]), r[1].strip()[:60]
        continue
    if r[2] not in ("-", ""):  # SASS row
        try:
            s = int(r[si])
        except ValueError:
            continue
        stalls = sorted(((int(r[i]) if r[i].strip().isdigit() else 0, h[6:]) for i, h in stall_cols), reverse=True)[:2]
        items.append((s, r[2], f"{fname}:{cur_line}", r[3].strip()[:70], stalls))
tot = sum(i[0] for i in items)
print("total", tot)
for s, addr, where, sass, stalls in sorted(items, reverse=True)[:n]:
    print(f"{s:6d} {100*s/tot:5.1f}% {addr[-6:]} {where:28s} {sass:70s} {stalls}")
