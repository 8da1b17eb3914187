"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list into per-kernel shares.
   python tools/launch_summary.py profiles/r01b_launches.csv"""
import csv, sys
from collections import defaultdict
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = defaultdict(float), defaultdict(int)
for r in rows[1:]:
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui].strip(), 1e-6)
    nm = r[ki].split("(")[0]
    tot[nm] += v * scale
    cnt[nm] += 1
T = sum(tot.values())
print(f"| kernel | launches | ms (serialised, cold cache) | share |\n|---|---|---|---|")
for k, v in sorted(tot.items(), key=lambda x: -x[1]):
    print(f"| `{k}` | {cnt[k]} | {v:.3f} | {100 * v / T:.1f}% |")
print(f"| total | {sum(cnt.values())} | {T:.3f} | 100% |")
