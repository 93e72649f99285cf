"""SASS opcode inventory of libpg_b200.so: counts of the Blackwell-native instructions (tcgen05.mma = UTCHMMA, TMA =
UTMALDG / UTMASTG / UTMAREDG, tcgen05.ld = LDTM, tcgen05.commit = UTCBAR, elect.sync = ELECT) per kernel.
    python tools/sass_inventory.py > profiles/r02_sass_inventory.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "pytorch_generative_b200", "libpg_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
ops, per, kern = collections.Counter(), collections.defaultdict(collections.Counter), None
WATCH = ("UTCHMMA", "UTMALDG", "UTMASTG", "UTMAREDG", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "SYNCS", "ELECT", "MUFU", "FENCE")
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and kern:
        op = m.group(1)
        base = op.split(".")[0]
        if base in WATCH:
            key = "UTCHMMA.2CTA" if op.startswith("UTCHMMA") and "2CTA" in op else base
            ops[key] += 1
            per[kern][key] += 1
print("# SASS opcode inventory of pytorch_generative_b200/libpg_b200.so (cuobjdump -sass, sm_100a)")
print("# tcgen05.mma = UTCHMMA (.2CTA = cta_group::2), TMA load / store / reduce = UTMALDG / UTMASTG / UTMAREDG,")
print("# tcgen05.ld = LDTM, tcgen05.commit = UTCBAR, mbarrier ops = SYNCS, elect.sync = ELECT")
for k, v in sorted(ops.items(), key=lambda kv: -kv[1]):
    print(f"{k:16s} {v}")
print("\n# per kernel: UTCHMMA(+2CTA) / UTMALDG / UTMASTG+REDG / LDTM / ELECT")
names = subprocess.run(["c++filt"], input="\n".join(per), capture_output=True, text=True).stdout.splitlines()
for raw, name in sorted(zip(per, names), key=lambda kn: kn[1]):
    c = per[raw]
    u = c["UTCHMMA"] + c["UTCHMMA.2CTA"]
    if u + c["UTMALDG"] == 0:
        continue
    short = re.sub(r"\(anonymous namespace\)::", "", name).split("(")[0].replace("void ", "")
    print(f"{u:4d} {c['UTMALDG']:4d} {c['UTMASTG'] + c['UTMAREDG']:4d} {c['LDTM']:4d} {c['ELECT']:4d}  {short}")
