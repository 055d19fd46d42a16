"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name.
usage: python tools/ncu_agg.py launches.csv [skip_first_n_launches]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[hi]
data = [r for r in rows[hi + 1:] if len(r) == len(hdr)][skip:]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for r in data:
    name = re.sub(r"\(.*", "", r[ki])
    name = re.sub(r".*::", "", name)[:44]
    v = float(r[vi].replace(",", ""))
    u = r[ui]
    v = v / 1e3 if u == "ns" else (v * 1e3 if u in ("ms", "msecond") else v)
    agg[name][0] += 1
    agg[name][1] += v
    tot += v
print(f"total {tot:.1f} us over {len(data)} launches")
for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:16]:
    print(f"{k:46s} {c:6d} {t:12.1f} us {100 * t / tot:5.1f}%  avg {t / c:9.1f}")
