"""Per-launch device times of the factorisation / Takahashi schedules (CUDA events, spde_plan_profile),
joined with the scheduled flops and tile counts of each launch.  Run on the GPU box:
    python tools/launch_profile.py c3 > profiles/rNN_launch_profile_c3.txt"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import torch

import bench
import plan_emulator as pe

name = sys.argv[1] if len(sys.argv) > 1 else "c3"
inp = bench.make_inputs(name)
mod = bench.build_ours(inp)
m = mod.mod
m.initFit(inp["data"], idx=inp["idx"])
m.logLike(inp["theta"], grad=True, exact_grad=True)        # warm
plan = m.engine.plan
plan.profile(True)
m.logLike(inp["theta"], grad=True, exact_grad=True)
torch.cuda.synchronize()
pms, pcnt = plan.profile(False)
kinds = ["gemm", "potrf", "extend_add", "memset", "gather", "wtw", "extract", "gemv"]
print("# per kind ms:", {k: round(float(pms[i].sum()), 2) for i, k in enumerate(kinds)})
print("# gemm by variant (cfg*4+ak*2+bk) ms:", {v: round(float(pms[0][v]), 2) for v in range(16) if pms[0][v] > 0})
for prog, pname in ((0, "factor"), (3, "selinv")):
    P = pe.Program(plan, prog)
    ms = plan.export(prog, 7, "f4")
    g = P.gemm
    M_, N_ = g["M"].astype(np.float64), g["N"].astype(np.float64)
    # lower: the trapezoid on and below the diagonal of an M x N block (M >= N), not half of it
    fl = 2.0 * g["K"] * np.where(g["flags"] & pe.GF_LOWER, M_ * N_ - 0.5 * N_ * np.minimum(M_, N_), M_ * N_)
    rows = []
    for i, L in enumerate(P.launches):
        if L["kind"] != 0:
            continue
        t = g[L["task0"]:L["task0"] + L["ntasks"]]
        f = fl[L["task0"]:L["task0"] + L["ntasks"]].sum()
        rows.append((ms[i], f, L["ntiles"], L["ntasks"], L["variant"], int(t["M"].max()), int(t["N"].max()), int(t["K"].max())))
    rows = np.array(rows)
    tot_ms, tot_f = rows[:, 0].sum(), rows[:, 1].sum()
    print("\n## %s: %d gemm launches, %.1f ms, %.3g flop -> %.2f TFLOP/s (last profiled run of this schedule)"
          % (pname, len(rows), tot_ms, tot_f, tot_f / tot_ms / 1e9))
    # bucket by achieved rate
    order = np.argsort(-rows[:, 0])
    print("top 25 launches by time:  ms   TFLOP/s  tiles tasks variant  maxM  maxN  maxK")
    for r in rows[order[:25]]:
        print("   %8.3f %8.2f %7d %5d %5d %7d %5d %6d" % (r[0], r[1] / r[0] / 1e9, r[2], r[3], r[4], r[5], r[6], r[7]))
    for lab, msk in (("N<=64,K<=64", (rows[:, 6] <= 64) & (rows[:, 7] <= 64)), ("N<=64,K>64", (rows[:, 6] <= 64) & (rows[:, 7] > 64)),
                     ("N>64,K<=256", (rows[:, 6] > 64) & (rows[:, 7] <= 256)), ("N>64,K>256", (rows[:, 6] > 64) & (rows[:, 7] > 256))):
        if msk.any():
            print("   class %-12s launches %5d  ms %9.2f  flop %.3g  -> %.2f TFLOP/s" % (lab, msk.sum(), rows[msk, 0].sum(), rows[msk, 1].sum(),
                                                                                         rows[msk, 1].sum() / max(rows[msk, 0].sum(), 1e-9) / 1e9))
    # efficiency against machine fill: tiles per launch relative to the 592 resident CTAs (148 SMs x 4)
    for lo, hi in ((0, 148), (148, 592), (592, 1776), (1776, 5920), (5920, 1 << 30)):
        msk = (rows[:, 2] >= lo) & (rows[:, 2] < hi)
        if msk.any():
            print("   tiles in [%5d,%6s) launches %5d  ms %9.2f  flop %.3g  -> %.2f TFLOP/s   (mean K %.0f, mean tasks %.1f)" % (
                lo, "inf" if hi > 1 << 29 else str(hi), msk.sum(), rows[msk, 0].sum(), rows[msk, 1].sum(),
                rows[msk, 1].sum() / max(rows[msk, 0].sum(), 1e-9) / 1e9, rows[msk, 7].mean(), rows[msk, 3].mean()))
    other = [(kinds[L["kind"]], ms[i]) for i, L in enumerate(P.launches) if 0 < L["kind"] < len(kinds)]      # (not the sync records)
    agg = {}
    for k, v in other:
        agg[k] = agg.get(k, 0.0) + float(v)
    print("   non-gemm ms:", {k: round(v, 2) for k, v in agg.items()})

# --- triangular solves with one right-hand side (conditional mean)
eng = m.engine
x = torch.randn(eng.n, 1, dtype=torch.float64, device="cuda")
eng.solve(1, x.clone())
plan.profile(True)
eng.solve(1, x.clone())
torch.cuda.synchronize()
pms, pcnt = plan.profile(False)
print("\n## solve_A k=1: per kind ms", {k: round(float(pms[i].sum()), 3) for i, k in enumerate(kinds)},
      "launches", {k: int(pcnt[i].sum()) for i, k in enumerate(kinds)})
for prog, pname in ((1, "forward"), (2, "backward")):
    P = pe.Program(plan, prog, 1)
    ms = plan.export(prog, 7, "f4", 1)
    order = np.argsort(-ms)[:8]
    print("  %s: %d launches, %.2f ms; slowest:" % (pname, len(ms), ms.sum()),
          [(round(float(ms[i]), 3), int(P.launches[i]["kind"]), int(P.launches[i]["ntiles"]), int(P.launches[i]["ntasks"])) for i in order])
t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
for _ in range(3):
    eng.solve(1, x.clone())
torch.cuda.synchronize(); t0.record()
for _ in range(5):
    eng.solve(1, x.clone())
t1.record(); torch.cuda.synchronize()
print("  graph-replayed solve_A k=1: %.2f ms" % (t0.elapsed_time(t1) / 5))
