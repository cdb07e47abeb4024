"""ncu launch list (--metrics gpu__time_duration.sum --csv) -> per-kernel totals and shares (JSON)."""
import csv, json, re, sys
from collections import defaultdict

path, cmd = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else ""
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
tot = defaultdict(lambda: [0, 0.0])
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    name = re.sub(r"^void ", "", name).replace("vg::<unnamed>::", "").replace("vg::", "")
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    us = v / 1e3 if unit in ("ns", "nsecond") else v if unit in ("us", "usecond") else v * 1e3
    tot[name][0] += 1
    tot[name][1] += us
total = sum(v[1] for v in tot.values())
out = dict(command=cmd, note="cold-cache, serialised per-launch times: compare SHARES with bench.py's "
           "kernel_breakdown, not absolutes", total_us=round(total, 1),
           kernels=[dict(kernel=k, launches=n, total_us=round(us, 1), share=round(us / total, 4),
                         avg_us=round(us / n, 2))
                    for k, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1])])
print(json.dumps(out, indent=1))
