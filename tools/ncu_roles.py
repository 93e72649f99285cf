"""Splits the warp-stall samples of a warp-specialised kernel by code region.  Regions are delimited by marker
opcodes given on the command line as name=OPCODE (first occurrence starts the region):
    python tools/ncu_roles.py src.csv producer=UTMALDG mma=UTCHMMA drain=UTMAREDG softmax=MUFU.EX2"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
markers = [a.split("=") for a in sys.argv[2:]]
hdr = None
seen = {}
for r in rows:
    if r and r[0] == "Line No":
        hdr = r
        si = hdr.index("# Samples")
        stall_cols = [(i, h[6:]) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if hdr is None or len(r) <= si or r[2] in ("-", ""):
        continue
    try:
        addr = int(r[2], 16)
        s = int(r[si])
    except ValueError:
        continue
    if addr in seen:
        continue
    seen[addr] = (s, r[3].strip(), {h: (int(r[i]) if r[i].strip().isdigit() else 0) for i, h in stall_cols})
addrs = sorted(seen)
starts = []
for name, op in markers:
    for a in addrs:
        if op in seen[a][1]:
            starts.append((a, name))
            break
starts.sort()
print("instructions", len(addrs), "samples", sum(v[0] for v in seen.values()))
def region(a):
    cur = "prologue"
    for s, name in starts:
        if a >= s - 0x400:  # the role's branch + waits precede its first marker opcode a little
            cur = name
    return cur
agg = {}
for a in addrs:
    reg = region(a)
    d = agg.setdefault(reg, {"samples": 0, "stalls": {}, "top": []})
    d["samples"] += seen[a][0]
    for k, v in seen[a][2].items():
        d["stalls"][k] = d["stalls"].get(k, 0) + v
    d["top"].append((seen[a][0], hex(a)[-5:], seen[a][1][:60]))
for reg, d in agg.items():
    st = sorted(d["stalls"].items(), key=lambda kv: -kv[1])[:5]
    print(f"\n== {reg}: {d['samples']} samples; stalls {st}")
    for s, a, t in sorted(d["top"], reverse=True)[:12]:
        print(f"   {s:6d} {a} {t}")
