#!/usr/bin/env python
"""Sum of gpu__time_duration per kernel name from an `ncu --metrics gpu__time_duration.sum --csv` log (launch list).
usage: python tools/ncu_launch_sum.py launches.csv [skip_first_n_launches]"""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hdr = None
tot = defaultdict(lambda: [0, 0.0])
n = 0
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        try:
            v = float(r[hdr.index("Metric Value")].replace(",", ""))
        except ValueError:
            continue
        unit = r[hdr.index("Metric Unit")]
        v = v / 1e3 if unit in ("ns", "nsecond") else (v * 1e3 if unit in ("ms", "msecond") else v)   # -> us
        n += 1
        if n <= skip:
            continue
        name = r[hdr.index("Kernel Name")].split("(")[0].split("::")[-1]
        tot[name][0] += 1
        tot[name][1] += v
for name, (c, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{name:50s} {c:5d} launches {us / 1e3:10.3f} ms")
print(f"{'total':50s} {sum(c for c, _ in tot.values()):5d} launches {sum(u for _, u in tot.values()) / 1e3:10.3f} ms")
