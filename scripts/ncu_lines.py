"""Per-source-line instruction counts and stall samples of one kernel from an ncu report.

    python scripts/ncu_lines.py <report.ncu-rep> <library.so> <kernel substring> [top N]

`ncu --page source --csv` lists SASS instructions with their counters but not their source lines;
`nvdisasm -g` lists the same instructions with //## File ..., line N markers.  Both are in
program order, so zipping them attributes every counter to a CUDA source line."""
import csv, io, os, re, subprocess, sys, tempfile, collections

rep, lib, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 50
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
assert lines_of, "kernel not found in library"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
hdr = rows[hi]
ie, sm = hdr.index("Instructions Executed"), hdr.index("# Samples")
sass = [r for r in rows[hi + 1:] if len(r) > ie and r[0].startswith("0x")]
assert len(sass) == len(lines_of), (len(sass), len(lines_of))
inst, samp = collections.Counter(), collections.Counter()
for r, l in zip(sass, lines_of):
    inst[l] += int(r[ie]); samp[l] += int(r[sm])
ti, ts = sum(inst.values()), sum(samp.values())
src = {}
print(f"total warp instructions {ti}, stall samples {ts}")
for l, n in (samp.most_common(top) if os.environ.get('VG_BY_STALL') else inst.most_common(top)):
    n = inst[l]
    if l and l[0] not in src:
        p = os.path.join(os.path.dirname(os.path.abspath(lib)), "..", "csrc", l[0])
        src[l[0]] = open(p).read().splitlines() if os.path.exists(p) else []
    text = src[l[0]][l[1] - 1].strip()[:90] if l and l[1] - 1 < len(src.get(l[0], [])) else ""
    print(f"{100 * n / ti:5.2f}% inst {100 * samp[l] / ts:5.2f}% stall  {l[0] if l else '?'}:{l[1] if l else 0:<5} {text}")

# optional region summary: VG_REGIONS="name:lo-hi,name:lo-hi" (line ranges of the main .cu file)
reg = os.environ.get("VG_REGIONS")
if reg:
    main = max(set(l[0] for l in inst if l), key=lambda f: sum(n for l, n in inst.items() if l and l[0] == f))
    print("regions of", main)
    for item in reg.split(","):
        name, rng = item.split(":")
        lo, hi = map(int, rng.split("-"))
        n = sum(v for l, v in inst.items() if l and l[0] == main and lo <= l[1] <= hi)
        s = sum(v for l, v in samp.items() if l and l[0] == main and lo <= l[1] <= hi)
        print(f"  {name:<14} {100 * n / ti:5.1f}% inst ({n / 1e6:8.1f} M)  {100 * s / ts:5.1f}% stall samples")
    n = sum(v for l, v in inst.items() if not l or l[0] != main)
    print(f"  {'inlined hdrs':<14} {100 * n / ti:5.1f}% inst")
