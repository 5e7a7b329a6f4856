"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: time and share per kernel."""
import csv, re, sys
from collections import defaultdict
rows = []
with open(sys.argv[1]) as fh:
    lines = [l for l in fh if not l.startswith("==")]
rd = csv.DictReader(lines)
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    rows.append((name, ns, r.get("Grid Size", ""), r.get("Block Size", "")))
tot = sum(r[1] for r in rows)
agg = defaultdict(lambda: [0, 0.0])
for n, ns, *_ in rows:
    agg[n][0] += 1
    agg[n][1] += ns
print("| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
for n, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("| `%s` | %d | %.3f | %.1f %% |" % (n, c, ns / 1e6, 100 * ns / tot))
print("| **all** | %d | %.3f | 100 %% |" % (len(rows), tot / 1e6))
if len(sys.argv) > 2:
    print("\nTop launches:")
    for n, ns, g, b in sorted(rows, key=lambda r: -r[1])[: int(sys.argv[2])]:
        print("  %8.1f us  %s grid=%s block=%s" % (ns / 1e3, n, g, b))

# optional per-layer attribution: python tools/summarize_launches.py launches.csv N trace.json
if len(sys.argv) > 3:
    import json
    trace = json.load(open(sys.argv[3]))
    i = 0
    per = []
    for label, k in trace:
        if label.startswith("?") or k == 0:
            continue
        seg = rows[i:i + k]
        i += k
        per.append((sum(r[1] for r in seg), label, "+".join(re.sub(r".*::", "", r[0]).replace("void ", "")[:28] for r in seg)))
    print("\nPer call (aligned with the engine's launch trace; %d of %d launches consumed):" % (i, len(rows)))
    for ns, label, names in sorted(per, key=lambda t: -t[0])[: int(sys.argv[2]) * 2]:
        print("  %8.1f us  %-70s %s" % (ns / 1e3, label, names))
