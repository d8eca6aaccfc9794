"""Per-kernel summary of an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file X.csv ...`):
kernel, launches, total us, share.  Usage: python scripts/launch_summary.py X.csv "header comment" > profiles/launches_rN_summary.csv"""
import csv
import re
import sys
from collections import defaultdict

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
h = rows[0]
ik, iv, iu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
agg = defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    if len(r) <= iv:
        continue
    v = float(r[iv].replace(",", ""))
    v = v / 1e3 if r[iu] in ("nsecond", "ns") else (v * 1e3 if r[iu] in ("msecond", "ms") else v)
    name = re.sub(r"\(.*", "", r[ik]).replace("void ", "").replace("b200::", "")
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
n = sum(v[0] for v in agg.values())
if len(sys.argv) > 2:
    print("# " + sys.argv[2])
print(f"# kernels {n}, total {tot / 1e3:.1f} ms")
print("kernel,launches,total_us,share_pct")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"\"{k}\",{v[0]},{v[1]:.1f},{v[1] / tot * 100:.2f}")
