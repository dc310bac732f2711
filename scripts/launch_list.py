"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` log into kernel, launches, total_ms, share:
python scripts/launch_list.py gpurun_out/launches.csv > profiles/rNN_bench_launch_list.csv"""
import csv
import sys
from collections import defaultdict

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
ki, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = defaultdict(float), defaultdict(int)
for r in rows[1:]:
    if r[mi] != "gpu__time_duration.sum":
        continue
    v = float(r[vi].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1e-6)
    name = r[ki].split("(")[0]
    tot[name] += v
    cnt[name] += 1
T = sum(tot.values())
print(f"# total kernel time {T:.2f} ms over {sum(cnt.values())} launches")
print("kernel,launches,total_ms,share")
for k in sorted(tot, key=tot.get, reverse=True):
    print(f"{k},{cnt[k]},{tot[k]:.3f},{tot[k] / T:.4f}")
