"""Per-kernel totals of an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file LIST ...`):
   python tools/launch_shares.py LIST.csv [skip_launches] -> time, share of the listed launches, launch count and average per kernel."""
import csv
import re
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 10]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
tot = defaultdict(float)
cnt = defaultdict(int)
for r in rows[1 + skip:]:
    ms = float(r[vi].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "msecond": 1.0, "ms": 1.0, "nsecond": 1e-6}.get(r[ui], 1e-6)
    name = re.sub(r"\(.*", "", r[ki])
    tot[name] += ms
    cnt[name] += 1
T = sum(tot.values())
print(f"{sum(cnt.values())} launches, total {T:.1f} ms (cold-cache and serialised under ncu: shares, not absolute times)")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"{v:10.3f} ms {100 * v / T:5.1f}% x{cnt[k]:3d} avg {v / cnt[k]:8.3f}  {k}")
