"""Aggregates an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` capture by kernel:
launches, time, DRAM bytes; prints the step totals and writes the tcgen05-GEMM bytes-per-launch JSON bench.py reads.
usage: python tools/summarize_dram.py capture.csv steps [out.json]"""
import collections
import csv
import json
import re
import sys

path, steps = sys.argv[1], int(sys.argv[2])
lines = [l for l in open(path) if not l.startswith("==")]
rows = collections.defaultdict(dict)
for r in csv.DictReader(lines):
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    if r["Metric Name"] == "gpu__time_duration.sum":
        v = v / 1e3 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1e3)      # -> us
    else:
        v = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)                     # -> bytes
    name = re.sub(r"<unnamed>::", "", re.sub(r"\(.*", "", r["Kernel Name"]))
    rows[r["ID"]]["name"] = name
    rows[r["ID"]][r["Metric Name"]] = v
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for d in rows.values():
    a = agg[d["name"]]
    a[0] += 1
    a[1] += d.get("gpu__time_duration.sum", 0.0)
    a[2] += d.get("dram__bytes_read.sum", 0.0)
    a[3] += d.get("dram__bytes_write.sum", 0.0)
tot = [sum(a[i] for a in agg.values()) for i in range(4)]
print(f"{tot[0]} launches over {steps} steps: {tot[1] / 1e3 / steps:.2f} ms/step (cold-cache, serialised), DRAM "
      f"{tot[2] / 1e9 / steps:.2f} GB read + {tot[3] / 1e9 / steps:.2f} GB written = {(tot[2] + tot[3]) / 1e9 / steps:.2f} GB per step")
print(f"{'ms/step':>9} {'share':>6} {'n/step':>6} {'GB/step':>8} {'TB/s':>6}  kernel")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    gb = (a[2] + a[3]) / 1e9
    print(f"{a[1] / 1e3 / steps:9.3f} {100 * a[1] / tot[1]:5.1f}% {a[0] / steps:6.1f} {gb / steps:8.3f} {gb / (a[1] * 1e-6) / 1e3 if a[1] else 0:6.2f}  {k[:100]}")
g = [a for k, a in agg.items() if "dx_gemm_tc_kernel" in k]
if g and len(sys.argv) > 3:
    n = sum(a[0] for a in g)
    json.dump({"dram_bytes_per_launch": sum(a[2] + a[3] for a in g) / n, "launches": n, "steps": steps,
               "step_dram_bytes": (tot[2] + tot[3]) / steps, "step_dram_read_bytes": tot[2] / steps,
               "step_dram_write_bytes": tot[3] / steps, "source": path}, open(sys.argv[3], "w"), indent=1)
