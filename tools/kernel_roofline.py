"""Per-kernel roofline table of the hot path (run on the GPU box):

    python tools/kernel_roofline.py [c3] > profiles/rNN_kernel_roofline.txt      (JSON copy next to it)

HBM-bound kernels (assembly K2/K3, scatter K4, logdet K6, k=1 solves K7, reductions K8/K9, adjoint K11 and
the helper kernels of the factor / Takahashi schedules) are timed with CUDA events on the stream they are
launched on, L2 flushed between repetitions, and reported as ALGORITHMIC bytes / time against the measured
HBM copy bandwidth (MEASURED_PEAKS.json, 6546.9 GB/s on this pool).  The assembly kernels are additionally
timed on the 256x256x100 mesh (BASELINE configs[3]): the precision itself (2.25 GB) fits, only its factor
does not.  The dense kernels are reported in TFLOP/s against cuBLAS DGEMM measured in the same run.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import torch

import bench
import plan_emulator as pe
from spdepy_b200.engine import Engine, to_dev

F64 = torch.float64
try:
    HBM = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    HBM_SRC = "MEASURED_PEAKS.json"
except Exception:
    HBM, HBM_SRC = 6546.9, "fallback (MEASURED_PEAKS.json of this pool, 2026-10-17)"

_flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
rows = []


def timeit(fn, reps=7):
    fn()
    ts = []
    for _ in range(reps):
        _flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def hbm_row(mesh, kernel, what, nbytes, ms, launches=1):
    gbs = nbytes / (ms * 1e-3) / 1e9
    rows.append({"mesh": mesh, "kernel": kernel, "replaces": what, "bound": "hbm", "alg_bytes": int(nbytes), "ms": ms,
                 "launches": launches, "achieved_gbs": gbs, "peak_gbs": HBM, "frac": gbs / HBM})


def assembly_rows(tag, m, par):
    eng, g = m.engine, m.grid
    Ns, n = eng.Ns, eng.n
    st = m._assemble(par)
    p = st["p"]
    V, dt = st["V"], st["dt"]
    Hd = to_dev(st["H"])
    face = m.Hvar
    t = timeit(lambda: eng.ah_stencil(g.hx, g.hy, Hd, face))
    hbm_row(tag, "k_ah_stencil", "AH::AH (ccode/A*H_2D_b*.cpp)", 8 * Ns * ((16 if face else 0) + 9), t)
    if m.wkind is not None:
        wsd = to_dev(st["ws"])
        fw = m.wkind != "const"
        t = timeit(lambda: eng.aw_stencil(g.hx, g.hy, wsd, None, fw, 3, m.wkind == "var"))
        hbm_row(tag, "k_aw_stencil", "Aw::Aw (ccode/A*w_2D_b*.cpp)", 8 * Ns * ((4 if fw else 0) + 9), t)
        aw = eng.aw_stencil(g.hx, g.hy, wsd, None, fw, 3, m.wkind == "var")
    else:
        aw = None
    ah = eng.ah_stencil(g.hx, g.hy, Hd, face)
    kap = st["kappa"]
    t = timeit(lambda: eng.combine_A(m.aflav, V, dt, kap, ah, aw))
    hbm_row(tag, "k_combine_A", "A = Dv + (DvDk - Ah + Aw) dt (SciPy adds)", 8 * Ns * (9 * (2 if aw is not None else 1) + 9 + (1 if m.kvar else 0)), t)
    A9 = st["A9"]
    t = timeit(lambda: eng.atda(A9, kap, V, 1))
    hbm_row(tag, "k_atda", "A.T@iDv@Qs@iDv@A (SciPy SpGEMM)", 8 * Ns * (9 + 25 + (1 if m.kvar else 0)), t)
    AtDA, Q0 = st["AtDA"], st["mod0"]["Q"]
    t = timeit(lambda: eng.fill_spacetime(AtDA, A9, kap, V, Q0, st["sigma"], dt, m.divide))
    hbm_row(tag, "k_fill_spacetime", "sparse.bmat block-tridiagonal stacking", 8 * (43 * n + (25 + 25 + 9) * Ns), t)
    return st


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c3"
    inp = bench.make_inputs(name)
    mod = bench.build_ours(inp)
    m = mod.mod
    m.initFit(inp["data"], idx=inp["idx"])
    par = inp["theta"]
    tag = "%dx%dx%d" % (inp["M"], inp["N"], inp["T"])
    eng = m.engine
    n, Ns = eng.n, eng.Ns
    m.logLike(par, grad=True, exact_grad=True)
    st = assembly_rows(tag, m, par)
    Q = st["Q"]
    plan = eng.plan
    stats = plan.stats()
    r, nobs = m.r, m._obs["nobs"]
    obs, cnt = m._obs["nodes"], m._obs["cnt"]
    tau = float(np.exp(par[-1]))
    data = to_dev(m.data.reshape(nobs, r))

    # ---- reductions / gradient contraction on the full mesh
    x = torch.randn(n, 1, dtype=F64, device="cuda")
    t = timeit(lambda: eng.q_apply(Q, x))
    hbm_row(tag, "k_q_apply (k=1)", "Q_c@mu / Q@mu (SciPy SpMV)", 8 * (43 * n + 2 * n), t)
    W = torch.zeros(43 * n, dtype=F64, device="cuda")
    t = timeit(lambda: eng.sddmm(x, x, -0.5, W))
    hbm_row(tag, "k_sddmm (k=1, accumulate)", "mu_c * (dQ_i@mu_c) for all i", 8 * (2 * 43 * n + n), t)
    t = timeit(lambda: Engine.dot(W, Q))
    hbm_row(tag, "k_dot_partial+k_final", "sum(W .* Q) (log sigma trace)", 16 * 43 * n, t, 2)
    t = timeit(lambda: eng.assembly_adjoint(W, st["A9"], st["kappa"], st["V"], st["sigma"], st["dt"], True))
    hbm_row(tag, "k_adj_time_reduce+k_adj_cell", "the npar materialised dQ_i", 8 * (43 * n + (44 + 25 + 9 + 9 + 1) * Ns), t, 2)
    t = timeit(lambda: Engine.residual_ss(data, x, obs))
    hbm_row(tag, "k_resid_partial+k_final", "sum((y - S mu)^2)", 8 * (2 * nobs * r) + 8 * nobs, t, 2)

    # ---- factor store: scatter, logdet, k=1 solve
    nlow = int((plan.export(4, 1, "i8") >= 0).sum())
    ncand = plan.export(4, 2, "i4").size
    eng.factorize(1, Q, cnt, tau)
    plan.profile(True)
    eng.factorize(1, Q, cnt, tau)
    torch.cuda.synchronize()
    eng.selinv(1)
    torch.cuda.synchronize()
    plan.profile(False)
    t = timeit(lambda: eng.logdet(1))
    hbm_row(tag, "k_logdet_partial+k_sum_final", "Factor.logdet()", 16 * n, t, 2)
    b = eng.scatter_obs(data, obs, tau)
    t = timeit(lambda: eng.solve(1, b.clone()))
    P1, P2 = pe.Program(plan, 1, 1), pe.Program(plan, 2, 1)
    hbm_row(tag, "solve_A k=1 (k_gemv_grouped x%d, graph)" % (len(P1.launches) + len(P2.launches)), "Factor.solve_A(b), one column",
            16 * stats["nnzL"] + 32 * n, t, len(P1.launches) + len(P2.launches))

    # ---- the reference's default (Hutchinson) gradient: 100 probe columns (advection_diffusion2D.py:200-207)
    nh1 = 100
    Vp = (2.0 * torch.randint(0, 2, (n, nh1), device="cuda") - 1.0).to(F64)
    Xs = Vp.clone()
    eng.solve(1, Xs)
    t = timeit(lambda: eng.solve(1, Xs.copy_(Vp)), reps=3)
    rows.append({"mesh": tag, "kernel": "solve_A k=100 (grouped FP64 GEMM tiles, graph)", "bound": "tensor", "ms": t,
                 "tflops": 4.0 * stats["nnzL"] * nh1 / (t * 1e-3) / 1e12, "flops": 4.0 * stats["nnzL"] * nh1})
    Wh = torch.zeros(43 * n, dtype=F64, device="cuda")
    t = timeit(lambda: eng.sddmm(Xs, Vp, 0.005, Wh))
    hbm_row(tag, "k_sddmm (k=100, accumulate)", "V .* (dQ_i@Q^-1 V) for all i (npar SpMMs)", 8 * (2 * n * nh1 + 2 * 43 * n), t)
    t = timeit(lambda: Engine.wdot(Xs, Vp, cnt))
    hbm_row(tag, "k_wdot_partial+k_final (k=100)", "tau trace sum(V .* S^T S Q_c^-1 V)", 8 * (2 * n * nh1 + n), t, 2)
    del Vp, Xs, Wh

    # ---- helper kernels of the schedules, from the per-launch profile
    kinds = ["gemm", "potrf", "extend_add", "memset", "gather", "wtw", "extract", "gemv"]
    for prog, pname in ((0, "factor"), (3, "selinv")):
        P = pe.Program(plan, prog)
        ms = plan.export(prog, 7, "f4")
        agg = {}
        for i, L in enumerate(P.launches):
            k = int(L["kind"])
            a = agg.setdefault(k, {"ms": 0.0, "bytes": 0.0, "flops": 0.0, "n": 0})
            a["ms"] += float(ms[i])
            a["n"] += 1
            if k == pe.LK_ZERO:
                a["bytes"] += 8.0 * (int(L["a1"]) - int(L["a0"]))
            elif k == pe.LK_EXTADD:
                t_ = P.ext[L["task0"]:L["task0"] + L["ntasks"]]
                a["bytes"] += float((t_["nr"].astype(np.float64) ** 2).sum()) * 0.5 * 24     # lower triangle: read + read-modify-write
            elif k == pe.LK_GATHER:
                t_ = P.gather[L["task0"]:L["task0"] + L["ntasks"]]
                a["bytes"] += float((t_["nr"].astype(np.float64) ** 2).sum()) * 16
            elif k == pe.LK_POTRF:
                t_ = P.potrf[L["task0"]:L["task0"] + L["ntasks"]]
                a["bytes"] += float((t_["b"].astype(np.float64) ** 2).sum()) * 24
                a["flops"] += float((t_["b"].astype(np.float64) ** 3).sum()) * (1.0 / 3 + 1.0 / 3)
            elif k == pe.LK_WTW:
                t_ = P.wtw[L["task0"]:L["task0"] + L["ntasks"]]
                a["bytes"] += float((t_["b"].astype(np.float64) ** 2).sum()) * 16
            elif k == pe.LK_EXTRACT:
                a["bytes"] += (int(L["a1"]) - int(L["a0"])) * (32 + 8 + 16.0)
            elif k == pe.LK_GEMM:
                g = P.gemm[L["task0"]:L["task0"] + L["ntasks"]]
                a["flops"] += float((2.0 * g["M"] * g["N"] * g["K"]).sum())
        for k, a in sorted(agg.items()):
            if k == pe.LK_GEMM or k >= 8:        # (lane fork / join records are not launches)
                continue
            kn = {1: "k_potrf (64x64 + inverse)", 2: "k_extend_add", 3: "cudaMemsetAsync (L store / arenas)", 4: "k_selinv_gather",
                  5: "k_wtw", 6: "k_extract", 7: "k_gemv_grouped"}[k]
            hbm_row(tag, "%s [%s schedule]" % (kn, pname), "CHOLMOD numeric / INLA qinv internals", a["bytes"], a["ms"], a["n"])
    # k_scatter_q is launched ahead of the schedule
    t = timeit(lambda: eng.factorize(1, Q, cnt, tau))
    rows.append({"mesh": tag, "kernel": "whole factor schedule (graph)", "bound": "tensor", "ms": t,
                 "tflops": stats["flops"] / (t * 1e-3) / 1e12, "flops": stats["flops"]})
    t = timeit(lambda: eng.selinv(1))
    rows.append({"mesh": tag, "kernel": "whole Takahashi schedule (graph)", "bound": "tensor", "ms": t,
                 "tflops": 2 * stats["flops"] / (t * 1e-3) / 1e12, "flops": 2 * stats["flops"]})

    # ---- assembly on the headline mesh (its factor does not fit; the precision does)
    del W, x, b
    torch.cuda.empty_cache()
    if name == "c3":
        import spdepy_b200 as sp
        M4, N4, T4 = 256, 256, 100
        xs, ys = np.linspace(0, 15 * (M4 - 1) / 49, M4), np.linspace(0, 15 * (N4 - 1) / 49, N4)
        ts = np.linspace(0, 2 * (T4 - 1) / 19, T4)
        g0 = sp.grid(x=xs, y=ys)
        m0 = sp.model(grid=g0, spde="whittle-matern", ha=False, anisotropic=False, bc=3, parameters=np.array([-2.0, -0.5, np.log(10.0)]))
        g4 = sp.grid(x=xs, y=ys, t=ts)
        m4 = sp.model(grid=g4, spde="advection-diffusion", ha=False, anisotropic=True, bc=3, mod0=m0).mod
        th = np.array([-1, -1, 1, -1, 1, -1, 0, np.log(1000.0)], dtype="float64")
        st4 = assembly_rows("256x256x100", m4, th)
        n4 = m4.engine.n
        x4 = torch.randn(n4, 1, dtype=F64, device="cuda")
        t = timeit(lambda: m4.engine.q_apply(st4["Q"], x4))
        hbm_row("256x256x100", "k_q_apply (k=1)", "Q@mu (SciPy SpMV)", 8 * (43 * n4 + 2 * n4), t)

    print("# HBM peak %.1f GB/s (%s); L2 flushed between repetitions; median of 7" % (HBM, HBM_SRC))
    print("%-12s %-52s %10s %9s %9s %7s %6s" % ("mesh", "kernel", "alg MB", "ms", "GB/s", "frac", "launch"))
    for r_ in rows:
        if r_["bound"] == "hbm":
            print("%-12s %-52s %10.2f %9.4f %9.1f %7.3f %6d" % (r_["mesh"], r_["kernel"], r_["alg_bytes"] / 1e6, r_["ms"],
                                                              r_["achieved_gbs"], r_["frac"], r_["launches"]))
        else:
            print("%-12s %-52s %10s %9.3f %9s %7s   (%.2f TFLOP/s on %.3g flop)" % (r_["mesh"], r_["kernel"], "-", r_["ms"], "-", "-",
                                                                                   r_["tflops"], r_["flops"]))
    out = os.path.join(ROOT, "gpurun_out", "kernel_roofline_%s.json" % name)
    os.makedirs(os.path.dirname(out), exist_ok=True)
    json.dump(rows, open(out, "w"), indent=1)


if __name__ == "__main__":
    main()
