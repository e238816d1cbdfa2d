"""Groups an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name: count, total, share."""
import csv, re, sys
from collections import defaultdict

rows = []
with open(sys.argv[1], newline="") as fh:
    lines = [l for l in fh if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
    name = re.sub(r"\(.*", "", r["Kernel Name"]).strip()
    rows.append((name, ns))
agg = defaultdict(lambda: [0, 0.0])
for n, t in rows:
    agg[n][0] += 1
    agg[n][1] += t
tot = sum(t for _, t in rows)
print(f"launches {len(rows)}  total {tot/1e6:.2f} ms (serialised, cold-cache: compare SHARES, not absolutes)")
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t/tot*100:6.2f}%  {t/1e6:10.3f} ms  {c:7d} x  avg {t/c/1e3:9.2f} us  {n}")
