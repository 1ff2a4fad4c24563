"""Sum DRAM traffic and time of all dense-tile launches (k_gemm_ws, k_gemm_grouped) of one evaluation from an ncu csv log
(ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:k_gemm --csv --log-file X.csv ...).
Writes profiles-style JSON to stdout."""
import csv, json, re, sys
path = sys.argv[1]
keys = ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum")
tot = {k: 0.0 for k in keys}
per = {}
ids = set()
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3}
with open(path, newline="") as f:
    rd = csv.reader(l for l in f if l.startswith('"'))
    hdr = next(rd)
    ii, mi, ui, vi, ki = hdr.index("ID"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value"), hdr.index("Kernel Name")
    for r in rd:
        if r[mi] in tot:
            v = float(r[vi].replace(",", "")) * scale.get(r[ui], 1.0)
            tot[r[mi]] += v
            name = re.sub(r"[<(].*", "", r[ki]).strip()
            a = per.setdefault(name, {k: 0.0 for k in keys + ("launches",)})
            a[r[mi]] += v
            if (name, r[ii]) not in ids:
                a["launches"] += 1
            ids.add((name, r[ii]))
n = len(ids)
out = {"kernel": "k_gemm_ws + k_gemm_grouped (all variants)", "launches": n,
       "dram_read_bytes": tot["dram__bytes_read.sum"], "dram_write_bytes": tot["dram__bytes_write.sum"],
       "traffic_bytes_per_launch": (tot["dram__bytes_read.sum"] + tot["dram__bytes_write.sum"]) / max(n, 1),
       "time_ms_under_ncu": tot["gpu__time_duration.sum"] * 1e3,
       "per_kernel": {k: {"launches": int(a["launches"]), "dram_read_bytes": a[keys[0]], "dram_write_bytes": a[keys[1]],
                          "time_ms_under_ncu": a[keys[2]] * 1e3} for k, a in per.items()},
       "how": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --profile-from-start off "
              "-k regex:k_gemm python tools/ncu_hbm_target.py c3 (one logLike(grad=True, exact_grad=True), plain launches)"}
print(json.dumps(out, indent=1))
