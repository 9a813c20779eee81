"""Transposed extract of `ncu -i X.ncu-rep --page raw --csv`: one row per metric whose name matches any of the
patterns, one column per captured kernel.  Usage: ncu_extract.py raw.csv out.csv [pattern ...]"""
import csv
import re
import sys

DEFAULT = [r"gpu__time_duration\.sum$", r"dram__bytes_(read|write)\.sum", r"dram__throughput", r"lts__t_bytes\.sum$",
           r"lts__throughput", r"l1tex__throughput", r"sm__throughput", r"sm__cycles_active\.avg$",
           r"dmma", r"pipe_tensor", r"pipe_fp64", r"sm__warps_active", r"launch__(registers|grid|block|occupancy)",
           r"smsp__issue_active", r"smsp__inst_executed\.sum$", r"smsp__warp_issue_stalled.*_per_warp_active\.pct$",
           r"smsp__average_warps?_issue_stalled.*per_issue_active"]

src, dst = sys.argv[1], sys.argv[2]
pats = [re.compile(p) for p in (sys.argv[3:] or DEFAULT)]
rows = list(csv.reader(open(src)))
hdr, units, data = rows[0], rows[1], rows[2:]
ki = hdr.index("Kernel Name")
with open(dst, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["metric", "unit"] + [r[ki][-48:] for r in data])
    for j, h in enumerate(hdr):
        if h == "Kernel Name" or any(p.search(h) for p in pats):
            w.writerow([h, units[j]] + [r[j] for r in data])
