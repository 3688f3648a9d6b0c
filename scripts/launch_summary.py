"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: launches, total us, share.
usage: launch_summary.py <launches.csv> [first_id last_id]"""
import collections
import csv
import re
import sys

path = sys.argv[1]
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 60
rows = []
with open(path) as f:
    for r in csv.reader(f):
        if len(r) < 15 or r[0] == "ID" or not r[0].isdigit():
            continue
        if r[12] != "gpu__time_duration.sum":
            continue
        i = int(r[0])
        if lo <= i <= hi:
            ns = float(r[14].replace(",", ""))
            if r[13] == "us":
                ns *= 1e3
            rows.append((i, r[4], r[8], ns))
agg = collections.OrderedDict()
for i, name, grid, ns in rows:
    short = re.sub(r"\(.*", "", name)
    short = re.sub(r"^void\s+", "", short)
    short = re.sub(r"agpt::|<unnamed>::|\(anonymous namespace\)::", "", short)
    a = agg.setdefault(short, [0, 0.0])
    a[0] += 1
    a[1] += ns
tot = sum(a[1] for a in agg.values())
print(f"{len(rows)} launches, {tot / 1e3:.1f} us of kernel time (ids {rows[0][0]}..{rows[-1][0]})")
for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"  {k:60s} n={n:5d}  {ns / 1e3:10.1f} us  {100 * ns / tot:5.1f}%  avg {ns / n / 1e3:7.1f} us")
