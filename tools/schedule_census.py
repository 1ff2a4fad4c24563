"""CPU-only census of the launch schedules of a mesh (no GPU needed): launches, tasks, tiles and flops per launch kind
and per tree level for the factorisation, the Takahashi pass and the k = 1 solve -- where the launch-latency-bound part of
a step lives.  Usage: python tools/schedule_census.py M N T [bc]   (e.g. 50 50 20 for BASELINE configs[1])"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np

import plan_emulator as pe
from spdepy_b200 import _lib

KINDS = ["gemm", "potrf", "extend_add", "memset", "gather", "wtw", "extract", "gemv", "sync", "copy"]


def census(plan, prog, k=0, name=""):
    P = pe.Program(plan, prog, k)
    L = P.launches
    g = P.gemm
    fl = 2.0 * g["K"] * np.where(g["flags"] & pe.GF_LOWER, g["M"].astype(float) * g["N"] - 0.5 * g["N"].astype(float) * np.minimum(g["M"], g["N"]),
                                 g["M"].astype(float) * g["N"])
    print("## %s: %d launches" % (name, len(L)))
    print("   kind         launches     tasks        tiles   flops")
    for kd in range(len(KINDS)):
        m = L["kind"] == kd
        if not m.any():
            continue
        f = 0.0
        if kd in (0, 7):
            for l in L[m]:
                f += fl[l["task0"]:l["task0"] + l["ntasks"]].sum()
        print("   %-11s %9d %9d %12d   %.3e" % (KINDS[kd], m.sum(), L["ntasks"][m].sum(), L["ntiles"][m].sum(), f))
    gm = L[L["kind"] == 0]
    if len(gm):
        edges = [0, 37, 148, 592, 2368, 1 << 62]
        print("   GEMM launches by tile count (148 SMs x 3-4 resident 64x64 tiles = 444-592 tiles per wave):")
        for a, b in zip(edges[:-1], edges[1:]):
            m = (gm["ntiles"] >= a) & (gm["ntiles"] < b)
            f = sum(fl[l["task0"]:l["task0"] + l["ntasks"]].sum() for l in gm[m])
            print("     tiles in [%5d, %s): %6d launches, %5.1f %% of the GEMM flops" % (a, "inf" if b > 1 << 60 else "%5d" % b, m.sum(),
                                                                                        100 * f / max(fl.sum(), 1)))
    return len(L)


if __name__ == "__main__":
    M, N, T = (int(v) for v in sys.argv[1:4])
    bc = int(sys.argv[4]) if len(sys.argv) > 4 else 3
    plan = _lib.PlanHandle(M, N, T, bc)
    st = plan.stats()
    print("# %dx%dx%d bc%d: n %d, %d supernodes, %d levels, nnz(L) %.3e, sum cc^2 %.3e, largest front %d" %
          (M, N, T, bc, st["n"], st["nsuper"], st["levels"], st["nnzL"], st["flops"], st["max_front"]))
    tot = census(plan, 0, name="factorisation")
    tot += census(plan, 3, name="Takahashi selected inverse")
    tot += census(plan, 1, 1, name="forward solve k=1") + census(plan, 2, 1, name="back solve k=1")
    print("# launches of one logLike + exact gradient (3-D posterior part): %d; at ~4 us of launch + drain latency each that is "
          "%.1f ms of pure latency" % (tot, tot * 4e-3))
