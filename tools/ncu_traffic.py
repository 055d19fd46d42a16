"""Sums dram bytes / durations of the kernels matching a name pattern in an `ncu --csv` metric log and writes the JSON that
bench.py reads for `roofline.traffic`.
usage: python tools/ncu_traffic.py log.csv name_regex skip_launches algorithmic_flops out.json"""
import csv
import json
import re
import sys

log, pat, skip, flops, out = sys.argv[1], re.compile(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4]), sys.argv[5]
rows = list(csv.reader(open(log, errors="ignore")))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[hi]
ki, mi, vi, ui, ii = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("ID")
per = {}
for r in rows[hi + 1:]:
    if len(r) != len(hdr) or not pat.search(r[ki]):
        continue
    v = float(r[vi].replace(",", ""))
    u = r[ui].lower()
    if r[mi].startswith("dram__bytes"):
        v *= {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
    elif r[mi].startswith("gpu__time"):
        v *= {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1, "msecond": 1}.get(u, 1e-6)
    per.setdefault(int(r[ii]), {})[r[mi]] = v
ids = sorted(per)[skip:]
rd = sum(per[i].get("dram__bytes_read.sum", 0) for i in ids)
wr = sum(per[i].get("dram__bytes_write.sum", 0) for i in ids)
ms = sum(per[i].get("gpu__time_duration.sum", 0) for i in ids)
res = {"source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none "
                 "python tools/factor_once.py chol 16384 16384 warm=1 lookahead=0 (second step)", "kernel_regex": sys.argv[2],
       "launches": len(ids), "dram_read_bytes": rd, "dram_write_bytes": wr,
       "traffic_bytes_per_launch": (rd + wr) / max(len(ids), 1), "kernel_ms_under_ncu": ms, "algorithmic_flops": flops}
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res))
