"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv --log-file X.csv` launch list per kernel name.
Usage: python tools/ncu_launch_summary.py X.csv "<command line that was profiled>" > profiles/rNN_ncu_launches_summary.txt"""
import csv, re, sys
from collections import OrderedDict
path = sys.argv[1]
cmd = sys.argv[2] if len(sys.argv) > 2 else ""
agg = OrderedDict()
tot = 0.0
n = 0
with open(path, newline="") as f:
    rows = (l for l in f if l.startswith('"'))
    rd = csv.reader(rows)
    hdr = next(rd)
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    for r in rd:
        name = re.sub(r"\(.*", "", r[ki]).strip()
        v = float(r[vi].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0}.get(r[ui], 1e-6)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1; a[1] += v
        tot += v; n += 1
print("# ncu --metrics gpu__time_duration.sum --clock-control none --csv %s" % cmd)
print("# per-launch times are cold-cache and serialised under ncu: compare SHARES with bench.py's roofline.gemm_share_of_scheduled_time")
gem = sum(v[1] for k, v in agg.items() if "k_gemm" in k)
print("# total kernels %d, total %.2f ms; dense tiles (k_gemm_ws + k_gemm_grouped, all variants) share = %.1f%%" % (n, tot, 100 * gem / max(tot, 1e-9)))
for k, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-95s n=%6d ms=%10.3f share=%5.1f%% avg_us=%9.1f" % (k[:95], c, ms, 100 * ms / tot, 1e3 * ms / c))
