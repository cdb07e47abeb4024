"""Stall-reason totals of one kernel from an ncu report, overall and per source-line range.

    python scripts/ncu_stalls.py <report.ncu-rep> <library.so> <kernel substring> [name:lo-hi,...]

Same zipping of `ncu --page source --csv` with `nvdisasm -g` line markers as scripts/ncu_lines.py."""
import csv, io, os, re, subprocess, sys, tempfile, collections

rep, lib, kern = sys.argv[1:4]
regions = []
if len(sys.argv) > 4:
    for item in sys.argv[4].split(","):
        name, rng = item.split(":")
        lo, hi = map(int, rng.split("-"))
        regions.append((name, lo, hi))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
lines_of = None
for f in sorted(os.listdir(tmp)):
    txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    m = re.search(r"\.text\.(\S*%s\S*):" % re.escape(kern), txt)
    if not m:
        continue
    body = txt[m.end():]
    end = body.find("\n//--------------------- .text.")
    body = body[:end] if end > 0 else body
    cur, lines_of = None, []
    for ln in body.splitlines():
        mm = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if mm:
            cur = (os.path.basename(mm.group(1)), int(mm.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
            lines_of.append(cur)
    break
assert lines_of, "kernel not found"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi_ = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
hdr = rows[hi_]
cols = [(i, c) for i, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
ie = hdr.index("Instructions Executed")
sass = [r for r in rows[hi_ + 1:] if len(r) > ie and r[0].startswith("0x")]
assert len(sass) == len(lines_of)
main = collections.Counter(l[0] for l in lines_of if l).most_common(1)[0][0]


def summarise(name, pred):
    tot = collections.Counter()
    inst = 0
    for r, l in zip(sass, lines_of):
        if not pred(l):
            continue
        inst += int(r[ie])
        for i, c in cols:
            tot[c] += int(r[i])
    n = sum(tot.values())
    top = ", ".join(f"{c[6:]} {100 * v / max(n, 1):.0f}%" for c, v in tot.most_common(6))
    print(f"{name:<14} samples {n:7d}  inst {inst / 1e6:8.1f} M   {top}")
    return n


summarise("all", lambda l: True)
for name, lo, hi in regions:
    summarise(name, lambda l: l and l[0] == main and lo <= l[1] <= hi)
