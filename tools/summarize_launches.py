#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (template
arguments kept short): launches, total us, share, average us.

    python tools/summarize_launches.py gpurun_out/launches.csv > profiles/rN_launches_summary.txt
"""
import csv
import re
import sys
from collections import OrderedDict


def short(name: str) -> str:
    name = name.replace("<unnamed>::", "").replace("ace::", "")
    m = re.match(r"(?:void )?([\w:]+)(<.*>)?\(", name)
    if not m:
        return name[:60]
    base, targs = m.group(1), m.group(2) or ""
    epi = re.search(r"Epi\w+", targs)
    nums = re.match(r"<\(int\)(\d+), \(int\)(\d+)", targs)
    tag = ""
    if nums:
        tag = f"<{nums.group(1)},{nums.group(2)}" + (f",{epi.group(0)}>" if epi else ">")
    elif epi:
        tag = f"<{epi.group(0)}>"
    elif targs:
        tag = re.sub(r"\(int\)", "", targs)[:24]
    return base + tag


def main(path: str) -> None:
    rows = []
    with open(path) as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(lines):
        if r["Metric Name"] == "gpu__time_duration.sum":
            rows.append((short(r["Kernel Name"]), float(r["Metric Value"]) / 1e3))
    agg = OrderedDict()
    for k, us in rows:
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += us
    total = sum(a[1] for a in agg.values())
    print(f"total_us {total:.1f}  launches {len(rows)}")
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:56s} launches={n:4d} us={us:10.1f} share={100 * us / total:5.1f}% avg_us={us / n:8.1f}")


if __name__ == "__main__":
    main(sys.argv[1])
