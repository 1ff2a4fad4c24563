"""Summarise an ncu report (read with `ncu -i X.ncu-rep --page raw --csv`) per kernel name: launches, total time,
DRAM bytes, achieved DRAM GB/s, L2 hit rate, tensor-pipe activity.  Usage: python tools/ncu_summarise.py X.ncu-rep [peak_gbs]"""
import csv, io, subprocess, sys, re
from collections import OrderedDict
rep = sys.argv[1]
peak = float(sys.argv[2]) if len(sys.argv) > 2 else 6449.1
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rd = list(csv.reader(io.StringIO(txt)))
hdr, units, body = rd[0], rd[1], rd[2:]
col = {h: i for i, h in enumerate(hdr)}
def num(row, key, default=0.0):
    if key not in col: return default
    try: return float(row[col[key]].replace(",", ""))
    except Exception: return default
def scale(key):
    u = units[col[key]] if key in col else ""
    return {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0}.get(u, 1.0)
agg = OrderedDict()
for r in body:
    name = re.sub(r"\(.*", "", r[col["Kernel Name"]]).strip()
    a = agg.setdefault(name, {"n": 0, "t": 0.0, "rd": 0.0, "wr": 0.0, "hit": 0.0, "tensor": 0.0, "dram_pct": 0.0})
    t = num(r, "gpu__time_duration.sum") * scale("gpu__time_duration.sum")
    a["n"] += 1; a["t"] += t
    a["rd"] += num(r, "dram__bytes_read.sum") * scale("dram__bytes_read.sum")
    a["wr"] += num(r, "dram__bytes_write.sum") * scale("dram__bytes_write.sum")
    a["hit"] += num(r, "lts__t_sector_hit_rate.pct") * t
    a["tensor"] += num(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active") * t
    a["dram_pct"] += num(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed") * t
print("# %s ; per-launch times under ncu are cold-cache and serialised; peak %.1f GB/s" % (rep, peak))
print("%-44s %6s %10s %10s %10s %9s %7s %8s %8s" % ("kernel", "n", "ms", "rd MB", "wr MB", "GB/s", "frac", "L2hit%", "dram%"))
for k, a in agg.items():
    gbs = (a["rd"] + a["wr"]) / max(a["t"], 1e-12) / 1e9
    print("%-44s %6d %10.4f %10.2f %10.2f %9.1f %7.3f %8.1f %8.1f" % (k[:44], a["n"], a["t"] * 1e3, a["rd"] / 1e6, a["wr"] / 1e6, gbs, gbs / peak,
                                                             a["hit"] / max(a["t"], 1e-12), a["dram_pct"] / max(a["t"], 1e-12)))
