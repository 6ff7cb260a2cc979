"""Per-kernel summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, total us, share.
usage: python tools/ncu_launch_summary.py launches.csv [first_launch_id [last_launch_id]]"""
import csv
import re
import sys

path = sys.argv[1]
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 60
rows = []
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    i = int(r["ID"])
    if lo <= i <= hi:
        rows.append((i, r["Kernel Name"], float(r["Metric Value"].replace(",", "")) / 1e3))
agg = {}
for _, name, us in rows:
    short = re.sub(r"^void\s+", "", name)
    short = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", short)
    short = short.split("(")[0]
    a = agg.setdefault(short, [0, 0.0])
    a[0] += 1
    a[1] += us
tot = sum(v[1] for v in agg.values())
print(f"{len(rows)} launches, {tot / 1e3:.3f} ms")
for k, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{us:10.1f} us  {100 * us / tot:5.1f} %  x{c:<4d} {k[:150]}")
