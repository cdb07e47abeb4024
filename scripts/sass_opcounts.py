"""SASS evidence per kernel of a built library: counts of the Blackwell-native mnemonics
(B200_PROFILING.md, 'What proves a Blackwell-native kernel').

    python scripts/sass_opcounts.py vilgod_b200/lib/libvilgod_b200.so > profiles/r02_sass_opcounts.txt
"""
import collections, re, subprocess, sys

lib = sys.argv[1]
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
OPS = ["UTCHMMA.2CTA", "UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "SYNCS", "MUFU.EX2",
       "MUFU.TANH", "FFMA2", "ATOMS", "REDUX", "UBLKCP", "HMMA", "LDGSTS"]
per = collections.OrderedDict()
cur = None
for ln in txt.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        per[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if not m:
        continue
    op = m.group(1)
    per[cur]["_total"] += 1
    for o in OPS:
        if op == o or op.startswith(o + "."):
            per[cur][o] += 1
            if o == "UTCHMMA.2CTA":
                break
demangle = subprocess.run(["c++filt"], input="\n".join(per), capture_output=True, text=True).stdout.splitlines()
print(f"# cuobjdump -sass {lib}: instructions per kernel (static counts)")
print(f"# {'kernel':<100} " + " ".join(f"{o:>12}" for o in ["total"] + OPS))
tot = collections.Counter()
for (k, c), name in zip(per.items(), demangle):
    name = re.sub(r"vg::\(anonymous namespace\)::", "", name)
    name = re.sub(r"\((?:const )?(?:__grid_constant__ )?.*\)$", "", name)[:100]
    print(f"  {name:<100} " + " ".join(f"{c[o]:>12}" for o in ["_total"] + OPS))
    tot.update(c)
print(f"  {'ALL KERNELS':<100} " + " ".join(f"{tot[o]:>12}" for o in ["_total"] + OPS))
print("# UTCHMMA* = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA tensor load / store,")
print("# UBLKCP = cp.async.bulk (TMA without a tensor map: the projection's background stores),")
print("# HMMA = legacy mma.sync (must be 0: no first-generation kernel ships)")
