"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: count, total time, share."""
import csv, re, sys
from collections import defaultdict

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
r = csv.DictReader(lines)
tot = defaultdict(float); cnt = defaultdict(int)
for row in r:
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    name = re.sub(r"^.*::", "", name)
    v = float(row["Metric Value"].replace(",", ""))
    unit = row.get("Metric Unit", "ns")
    ns = v * {"ns": 1, "us": 1e3, "usecond": 1e3, "msecond": 1e6, "nsecond": 1, "ms": 1e6, "second": 1e9}.get(unit, 1)
    tot[name] += ns; cnt[name] += 1
total = sum(tot.values())
print(f"total kernel time {total / 1e6:.2f} ms over {sum(cnt.values())} launches")
print(f"{'kernel':60s} {'launches':>9s} {'total_ms':>10s} {'avg_us':>9s} {'share':>7s}")
for k in sorted(tot, key=tot.get, reverse=True):
    print(f"{k[:60]:60s} {cnt[k]:9d} {tot[k] / 1e6:10.3f} {tot[k] / cnt[k] / 1e3:9.2f} {100 * tot[k] / total:6.2f}%")
