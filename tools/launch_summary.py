"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name (shares of the step)."""
import collections, csv, re, sys

path = sys.argv[1]
rows = list(csv.reader(open(path, errors="ignore")))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr, data = rows[hi], rows[hi + 1:]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in data:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
    name = re.sub(r"\(.*", "", r[ki])
    name = re.sub(r"void |at::native::|\(anonymous namespace\)::", "", name)[:90]
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v for _, v in agg.values())
print(f"# {path}: {sum(c for c, _ in agg.values())} launches, {tot / 1e3:.2f} ms of kernel time (serialised, cold-cache: compare shares)")
print(f"{'share':>7} {'total ms':>10} {'n':>6} {'avg us':>9}  kernel")
for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print(f"{v / tot * 100:6.2f}% {v / 1e3:10.2f} {c:6d} {v / c:9.1f}  {k}")
