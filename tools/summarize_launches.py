"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import csv, re, sys, collections
path = sys.argv[1]
rows = []
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
r = csv.DictReader(lines)
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
n = 0
for row in r:
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = row["Kernel Name"]
    name = re.sub(r"\(.*", "", name)
    name = re.sub(r"<unnamed>::", "", name)
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    agg[name][0] += 1; agg[name][1] += us; tot += us; n += 1
print(f"{n} launches, {tot/1e3:.2f} ms total (cold-cache, serialised)")
for k, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{us/1e3:9.3f} ms {100*us/tot:5.1f}%  n={c:4d}  {k[:110]}")
