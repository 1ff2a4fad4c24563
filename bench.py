"""bench.py -- log-likelihood + exact-gradient evaluations per second of the SPDE hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c2|c1] [--impl ours|reference]

Metric (BASELINE.json): "loglik+grad evals/sec ... vs host-core reference".  One *step* is one
``logLike(theta, grad=True, exact_grad=True)`` of the workload's model: assemble Q, factorise
Q + tau S^T S (3-D) and the two 2-D matrices the prior's determinant collapses to, logdets, conditional
mean, quadratic forms, Takahashi selected inverses, gradient contraction.  Default workload: configs[2] of BASELINE.json (var-advection-
var-diffusion on the SINMOD-shaped 100x100x50 mesh, n = 5e5, 92 parameters) -- the largest config
whose FP64 factor fits one B200; the headline 256x256x100 mesh needs a 262 GB factor (SURVEY.md
finding 8) and is not a single-GPU configuration.

N > 1 (torchrun, one rank per GPU): weak scaling over independent theta evaluations, each rank
owning its own factorisation; the only collective is one NCCL all-reduce of the npar+1 likelihood /
gradient scalars per step (BASELINE.json north_star).

``--impl reference``: the reference's CPU path.  The reference's Python cannot travel to the GPU box
and its factoriser (CHOLMOD) is absent from the image, so this arm times the oracle port
(oracle/spde_oracle.py + oracle/cpu_cholesky.py: the reference's algorithm in SciPy + a supernodal
Cholesky on LAPACK, all host threads) on a bounded sample of the same workload and scales it.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "loglik+grad evals/sec"
UNIT = "evals/s"

WORKLOADS = {
    # name: (spde, mod0 spde, ha, ani, bc, M, N, T, description)
    "c1": ("whittle-matern", None, False, True, 3, 30, 30, None, "stationary Whittle-Matern 30x30 (configs[0])"),
    "c2": ("advection-diffusion", "whittle-matern", False, True, 3, 50, 50, 20,
           "advection-diffusion 50x50x20, 5000 obs (configs[1])"),
    "c3": ("var-advection-var-diffusion", "var-whittle-matern", False, True, 1, 100, 100, 50,
           "var-advection-var-diffusion 100x100x50 SINMOD-shaped, 92 parameters, 10% obs (configs[2])"),
    # batched sweep: every step is one theta-evaluation (likelihood + exact gradient) PLUS 1024 prior samples at that
    # theta (3-D factor of Q, 1024-column back substitution); thetas are sharded over the ranks
    "c5": ("var-advection-var-diffusion", "var-whittle-matern", False, True, 1, 100, 100, 50,
           "batched sweep on 100x100x50: per theta logLike+exact gradient and 1024 samples, thetas sharded over the GPUs (configs[4])"),
}
# half-resolution proxy of configs[3] (256x256x100, whose 260 GB FP64 factor does not fit one B200):
# same model, cell size, theta and 1 % observation density on 128x128x50 (n = 8.2e5, sum cc^2 = 1.57e13)
WORKLOADS["c4h"] = ("advection-diffusion", "whittle-matern", False, True, 3, 128, 128, 50,
                    "advection-diffusion 128x128x50, 1% obs: half-resolution proxy of configs[3] (256x256x100 does not fit)")
# configs[3] itself: streamed evaluation (spdepy_b200/csrc/ooc.cu) -- the factor never lives in HBM as a whole.  One
# evaluation takes minutes, so this workload is run explicitly (`--workload c4 --steps 1 --warmup 0`), never by default.
WORKLOADS["c4"] = ("advection-diffusion", "whittle-matern", False, True, 3, 256, 256, 100,
                   "advection-diffusion 256x256x100, 1% obs (configs[3]), streamed depth-first evaluation with pinned-host panel store")
N_SAMPLES = {"c5": 1024}


def make_inputs(name, M=None, N=None, T=None, seed=0):
    """Synthetic mesh, theta, observation indices and data for a workload (optionally on a smaller mesh
    with the same cell size, for the bounded CPU sample)."""
    spde, spde0, ha, ani, bc, M0, N0, T0, _ = WORKLOADS[name]
    M, N, T = M or M0, N or N0, (T or T0) if T0 else None
    if name in ("c3", "c5"):
        x, y, t = 800.0 * np.arange(M), 800.0 * np.arange(N), 10.0 * np.arange(T)
        theta = np.load(os.path.join(ROOT, "tests", "golden", "c3_theta.npy"))
        p0 = np.hstack([theta[55:91], theta[-1]])
        frac = 0.10
    elif name in ("c2", "c4h", "c4"):
        x, y, t = np.linspace(0, 15 * (M - 1) / 49, M), np.linspace(0, 15 * (N - 1) / 49, N), np.linspace(0, 2 * (T - 1) / 19, T)
        p0 = np.array([-2.0, -0.5, np.log(10.0)])
        theta = np.array([-1, -1, 1, -1, 1, -1, 0, -2, -0.5, np.log(1000.0)], dtype="float64")
        frac = 0.10 if name == "c2" else 0.01
    else:
        x = y = np.linspace(2 / 3, 40 - 2 / 3, M)
        t, p0 = None, None
        theta = np.array([-1, -1, 0.1, 0.1, np.log(100.0)])
        frac = 0.5
    n = M * N * (T or 1)
    rng = np.random.default_rng(5 + seed)
    idx = np.sort(rng.choice(n, int(frac * n), replace=False))
    data = rng.normal(size=(idx.size, 1))
    return dict(spde=spde, spde0=spde0, ha=ha, ani=ani, bc=bc, x=x, y=y, t=t, theta=theta, p0=p0, idx=idx, data=data,
                M=M, N=N, T=T, n=n, iso0=(name in ("c2", "c4h", "c4")))


def build_ours(inp):
    import spdepy_b200 as sp
    g = sp.grid(x=inp["x"], y=inp["y"], t=inp["t"])
    kw = {}
    if inp["t"] is not None:
        g0 = sp.grid(x=inp["x"], y=inp["y"])
        kw["mod0"] = sp.model(grid=g0, spde=inp["spde0"], ha=inp["ha"], anisotropic=(inp["ani"] and not inp["iso0"]),
                              bc=inp["bc"], parameters=inp["p0"])
    mod = sp.model(grid=g, spde=inp["spde"], ha=inp["ha"], anisotropic=inp["ani"], bc=inp["bc"], **kw)
    return mod


def build_oracle(inp):
    import spde_oracle as so
    import spdepy_b200 as sp                      # host-side grid classes only (input producers)
    g = sp.grid(x=inp["x"], y=inp["y"], t=inp["t"])
    names = {"whittle-matern": "whittle-matern-%s-2D", "var-whittle-matern": "var-whittle-matern-%s-2D"}
    if inp["t"] is None:
        return so.OracleSPDE(names[inp["spde"]] % "anisotropic", g, bc=inp["bc"]), g
    g0 = sp.grid(x=inp["x"], y=inp["y"])
    o0 = so.OracleSPDE(names[inp["spde0"]] % ("isotropic" if inp["iso0"] else "anisotropic"), g0, bc=inp["bc"], par=inp["p0"])
    return so.OracleSPDE(inp["spde"] + "-2D", g, mod0=o0, bc=inp["bc"]), g


# ---------------------------------------------------------------------------------------------
# the reference's CPU path (oracle port), timed on a bounded sample and scaled to the workload

def cpu_reference(name, full_stats, budget_s=25.0, nh1=100):
    import cpu_cholesky as cc
    import spde_oracle as so
    from spdepy_b200 import _lib
    spde, spde0, ha, ani, bc, M0, N0, T0, _ = WORKLOADS[name]
    if name in ("c3", "c5"):
        M, N, T = 24, 24, 10
    elif name in ("c2", "c4h", "c4"):
        M, N, T = 24, 24, 10
    else:
        M, N, T = M0, N0, None
    inp = make_inputs(name, M, N, T)
    mod, g = build_oracle(inp)
    plan = _lib.PlanHandle(g.shape[0], g.shape[1], T or 1, bc)
    mod.initFit(inp["data"], idx=inp["idx"])

    class _NoFactor:            # times the assembly alone (the reference factorises inside makeQ)
        def __init__(self, A, perm=None):
            pass

    def best_of(fn, reps):
        ts, out = [], None
        for _ in range(reps):
            t0 = time.perf_counter()
            out = fn()
            ts.append(time.perf_counter() - t0)
        return min(ts), out

    # every component is timed directly on the sample mesh (no differences of timings):
    so.set_factor(_NoFactor)
    try:
        t_make, (Q, _, dQ) = best_of(lambda: mod.makeQ(inp["theta"], grad=True), 2)      # assembly + the npar sparse dQ_i
    finally:
        so.set_factor(None, None)
    rngp = np.random.default_rng(4)
    V = (2 * rngp.integers(1, 3, plan.n * nh1) - 3).reshape(plan.n, nh1)
    mu = rngp.normal(size=(plan.n, 1))
    t_spmm, _ = best_of(lambda: [(d @ V, d @ mu) for d in dQ], 1)                       # advection_diffusion2D.py:204-206
    have_plan = Q.shape[0] == plan.n
    t_fac, fq = best_of(lambda: cc.SupernodalFactor(Q, plan=plan) if have_plan else so.DenseFactor(Q), 3)
    t_solveA, _ = best_of(lambda: fq.solve_A(V), 2)
    # one real evaluation of the port (value check + its own wall time, reported but not used for the scaling)
    so.set_factor(cc.factor_with_plan(plan))
    try:
        np.random.seed(4)
        t0 = time.perf_counter()
        like, jac = mod.logLike(inp["theta"], nh1=nh1, grad=True)
        t_eval = time.perf_counter() - t0
    finally:
        so.set_factor(None, None)
    s = plan.stats()
    # host dense rate (all threads): the factor / solve parts of the full workload are modelled at this
    # rate, which is optimistic for the CPU (a supernodal Cholesky never sustains the DGEMM rate)
    A = np.random.default_rng(0).normal(size=(2500, 2500))
    A @ A
    t0 = time.perf_counter()
    A @ A
    host_gflops = 2 * 2500 ** 3 / (time.perf_counter() - t0) / 1e9
    t_solve = 2.0 * t_solveA                                   # TrQ and TrQc (the r-column solve is negligible)
    t_asm = t_make + t_spmm                                    # assembly, dQ construction, SpMMs: O(n) sparse work
    full_asm = t_asm * full_stats["n"] / s["n"]
    full_fac = min(t_fac * full_stats["flops"] / s["flops"], full_stats["flops"] / (host_gflops * 1e9))
    full_solve = min(t_solve * full_stats["nnzL"] / s["nnzL"], 4.0 * full_stats["nnzL"] * 2 * nh1 / (host_gflops * 1e9))
    full_t = full_asm + 2 * full_fac + full_solve
    nsamp = N_SAMPLES.get(name, 0)
    if nsamp:      # one more factorisation is already counted (makeQ); add the nsamp-column back substitution
        full_t += 2.0 * full_stats["nnzL"] * nsamp / (host_gflops * 1e9)
    return {
        "value": 1.0 / full_t, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
        "sample": ("oracle port (SciPy assembly + supernodal Cholesky on LAPACK, Hutchinson nh1=%d as the reference does) of "
                   "one logLike(grad=True) on a %dx%dx%s mesh of the same model: %.2f s wall; components timed directly: assembly + "
                   "dQ list %.2f s, dQ SpMMs %.2f s, one factorisation %.2f s, two 100-column solves %.2f s. Scaled to the workload: "
                   "assembly and SpMMs by n; factor by sum cc^2 and solves by 4 nnz(L) k, each capped at the host's measured DGEMM "
                   "rate of %.0f GFLOP/s (optimistic for the CPU) -> %.1f s + 2 x %.1f s + %.1f s per evaluation"
                   % (nh1, M, N, T, t_eval, t_make, t_spmm, t_fac, t_solve, host_gflops, full_asm, full_fac, full_solve)),
        "sample_seconds": t_eval, "sample_like": float(like), "host_dgemm_gflops": host_gflops,
    }


# ---------------------------------------------------------------------------------------------

class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index),
                 "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
                 "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
                 "clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], 0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def fp64_peak():
    """cuBLAS DGEMM 8192^3 through torch.matmul: the FP64 roofline denominator (MEASURED_PEAKS.json has
    HBM and bf16 only).  Burst figure, best of 3."""
    import torch
    n = 8192
    A = torch.randn(n, n, dtype=torch.float64, device="cuda")
    B = torch.randn(n, n, dtype=torch.float64, device="cuda")
    torch.matmul(A, B)
    best = 1e30
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        torch.matmul(A, B)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    del A, B
    torch.cuda.empty_cache()
    return 2 * n ** 3 / best / 1e9     # TFLOP/s


def run_streamed(args):
    """configs[3] (256x256x100) on one B200: logLike + exact gradient with the streamed evaluator.  Prints the same
    JSON line as the in-core workloads; `roofline.achieved` is the whole-pass rate (algorithmic flops of the
    posterior factorisation + Takahashi pass / device time of the two passes), a lower bound of the GEMM rate."""
    import torch
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    from spdepy_b200 import _lib
    from spdepy_b200.engine import COUNTERS
    name = args.workload
    t0 = time.time()
    inp = make_inputs(name, args.mesh[0], args.mesh[1], args.mesh[2]) if args.mesh else make_inputs(name)
    mod = build_ours(inp)
    m = mod.mod
    m.initFit(inp["data"], idx=inp["idx"])
    eng = m.engine
    eng.streamed = True
    m.check_selinv = True
    plan = eng.plan
    stats = plan.stats()
    t_plan = time.time() - t0
    t0 = time.time()
    ooc = eng.ooc(True)
    ost = ooc.stats()
    t_ooc = time.time() - t0
    print("# plan %.1f s, streamed plan %.1f s: %s" % (t_plan, t_ooc, json.dumps(ost)), file=sys.stderr, flush=True)
    theta = inp["theta"]
    passes = []

    def step():
        like, jac = m.logLike(theta, grad=True, exact_grad=True)
        passes.append((ooc.info_d(1), ooc.info_d(2)))
        return like, jac

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    COUNTERS["h2d"] = COUNTERS["d2h"] = 0
    _lib.lib.spde_launch_count(1)
    with ClockSampler(0) as clk:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            like, jac = step()
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    launches = int(_lib.lib.spde_launch_count(0))
    fwd_ms, bwd_ms = passes[-1]
    print("# evaluation: %.1f s (forward %.1f s, backward %.1f s), like %.12g" % (ms / args.steps / 1e3, fwd_ms / 1e3, bwd_ms / 1e3, like),
          file=sys.stderr, flush=True)
    # full-size checks: residual of the conditional mean, Q_c mu = tau S^T y, and the gradient-free value (forward pass
    # only, quadratic form from |L^-1 P b|^2) against the value of the full evaluation
    tau = float(np.exp(theta[-1]))
    Q = m._state["Q"]
    mu = m.last["mu_c"]
    data_dev = torch.as_tensor(inp["data"], device="cuda")
    b = eng.scatter_obs(data_dev, m._obs["nodes"], tau)
    res = eng.q_apply(Q, mu) + mu * (m._obs["cnt"] * tau)[:, None] - b
    resid = float(res.abs().max() / b.abs().max())
    del res, b
    trace_check = m.last.get("selinv_trace_over_n")
    by_kind = None
    if args.profile_step:
        # one more evaluation with per-launch CUDA events (outside the timed region): device time per launch kind
        plan.profile(True)
        m.logLike(theta, grad=True, exact_grad=True)
        torch.cuda.synchronize()
        pms, pcnt = plan.profile(False)
        kinds = ["gemm", "potrf", "extend_add", "memset", "gather", "wtw", "extract", "gemv"]
        by_kind = {"ms": {k: float(pms[i].sum()) for i, k in enumerate(kinds)},
                   "launches": {k: int(pcnt[i].sum()) for i, k in enumerate(kinds)},
                   "passes_ms": [ooc.info_d(1), ooc.info_d(2)],
                   "gemm_tflops": (alg_all := 3.0 * stats["flops"] + ost["recompute_flops"]) / (float(pms[0].sum()) * 1e-3) / 1e12,
                   "note": "profiled evaluation (events around every launch); gemm_tflops = (3 sum cc^2 + flops factorised twice) / "
                           "summed k_gemm_grouped time; host<->device panel copies and update-matrix moves are not launches: "
                           "their time is passes_ms minus the sum over kinds"}
    cpu = None
    if not args.no_cpu:
        cpu = cpu_reference(name, stats)
        if name == "c4":
            cpu["sample"] += ("; at this size the CPU path cannot actually run: its two factors need 2 x 260 GB against %d GB "
                              "of host RAM, so the scaled figure is a time model, not a measurement" % (os.sysconf("SC_PAGE_SIZE") * os.sysconf("SC_PHYS_PAGES") // 10 ** 9))
    like_fwd = None
    if args.check_forward:
        like_fwd = float(m.logLike(theta, grad=False))
    alg = 3.0 * stats["flops"]
    peak = fp64_peak()
    achieved = alg / ((fwd_ms + bwd_ms) * 1e-3) / 1e12
    line = {
        "metric": METRIC, "value": args.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": WORKLOADS[name][8], "mesh": [inp["M"], inp["N"], inp["T"]], "n": inp["n"], "npar": int(theta.size),
                   "nobs": int(inp["idx"].size), "gradient": "exact (Takahashi selected inversion)",
                   "parallelism": "one GPU, streamed: %d segments (%d front-by-front), device pool %.1f GB, pinned host %.1f GB, "
                                  "%.3g flop factorised twice" % (ost["segments"], ost["top_segments"], ost["pool_bytes"] / 1e9,
                                                                   ost["host_bytes"] / 1e9, ost["recompute_flops"]),
                   "l2": "inputs larger than L2 (factor %.1f GB, streamed)" % (stats["factor_bytes"] / 1e9)},
        "cholesky_gflops": stats["flops"] / (fwd_ms * 1e-3) / 1e9,
        "forward_pass_ms": fwd_ms, "backward_pass_ms": bwd_ms,
        "symbolic": {k: stats[k] for k in ("nsuper", "nnzL", "flops", "factor_bytes", "levels", "max_front")},
        "streamed": ost, "host_setup_s": {"symbolic_and_plan": t_plan, "streamed_plan": t_ooc},
        "e2e": {"value": args.steps / (ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": COUNTERS["h2d"] // max(args.steps, 1),
                "d2h_bytes_per_step": COUNTERS["d2h"] // max(args.steps, 1),
                "note": "host inputs are copied inside the step (data, theta); the pinned-host panel traffic of the streamed "
                        "evaluator (%.1f GB each way per evaluation) is inside the timed region too" % (ost["host_bytes"] / 1e9)},
        "gpu_launches": launches, "clocks": clk.summary(),
        "roofline": {"bound": "tensor", "kernel": "k_gemm_grouped (FP64 DMMA m8n8k4)", "achieved": achieved, "peak": peak,
                     "unit": "TFLOP/s", "frac": achieved / peak, "traffic": None,
                     "note": "whole-pass rate: 3 sum cc^2 / (forward + backward device time), includes scatter, extend-add, "
                             "host transfers and the recomputed subtrees -- a lower bound of the kernel's own rate",
                     "peak_source": "cuBLAS DGEMM 8192^3 measured in this run", "algorithmic_flops_per_step": alg},
        "cpu_baseline": cpu,
        "checks": {"conditional_mean_residual_inf": resid, "like_full": float(like), "like_forward_only": like_fwd,
                   "grad_inf_norm": float(np.abs(jac).max()), "selinv_trace_over_n": trace_check},
        "profile_by_kind": by_kind, "passes_ms_per_step": [list(p) for p in passes],
        "last_like": float(like), "last_jac": [float(v) for v in jac],
    }
    print(json.dumps(line))
    return line


def run_ours(args):
    if args.workload == "c4" or args.streamed:
        return run_streamed(args)
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from spdepy_b200 import _lib
    from spdepy_b200.engine import COUNTERS, to_dev

    name = args.workload
    inp = make_inputs(name)
    mod = build_ours(inp)
    m = mod.mod
    m.initFit(inp["data"], idx=inp["idx"])
    # independent theta per rank and step (the optimiser's line search / FD checks of the reference)
    rng = np.random.default_rng(7 + rank)
    thetas = [inp["theta"] + 0.01 * rng.normal(size=inp["theta"].size) * (world > 1) for _ in range(args.steps + args.warmup)]
    plan = m.engine.plan
    stats = plan.stats()
    npar = inp["theta"].size
    red = torch.zeros(npar + 1, dtype=torch.float64, device="cuda")

    nsamp = N_SAMPLES.get(name, 0)
    gen = torch.Generator(device="cuda").manual_seed(8 + rank)
    sample_chk = [None]

    def step(theta, resident):
        if resident:
            if "data" not in m._obs:
                m._obs["data"] = to_dev(inp["data"])
        else:
            m._obs.pop("data", None)
            m.data = pinned_data.numpy()
        like, jac = m.logLike(theta, grad=True, exact_grad=True)
        if nsamp:
            # Model.sample at this theta (model.py:73-87): x = P^T L^-T z with L the factor of the 3-D prior
            eng = m.engine
            eng.factorize(0, m._state["Q"])
            chk = 0.0
            for c0 in range(0, nsamp, 512):
                z = torch.randn(eng.n, min(512, nsamp - c0), dtype=torch.float64, device="cuda", generator=gen)
                x = eng.solve(0, z, 10)
                chk += float((x * x).sum())
            sample_chk[0] = chk / (eng.n * nsamp)
        if world > 1:
            red[0] = like
            red[1:] = torch.as_tensor(jac, device="cuda")
            dist.all_reduce(red)
        return like, jac

    pinned_data = torch.from_numpy(inp["data"].copy()).pin_memory()

    def timed_region(resident):
        for k in range(args.warmup):
            step(thetas[k], resident)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        COUNTERS["h2d"] = COUNTERS["d2h"] = 0
        _lib.lib.spde_launch_count(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        out = None
        for k in range(args.steps):
            out = step(thetas[args.warmup + k], resident)
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            tms = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
            ms = float(tms.item())
        return ms, int(_lib.lib.spde_launch_count(0)), dict(COUNTERS), out

    with ClockSampler(local) as clk:
        ms, launches, _, last = timed_region(resident=True)
    clocks = clk.summary()
    ms_e2e, _, cnt_e2e, _ = timed_region(resident=False)

    line = None
    if rank == 0:
        # one profiled step outside the timed region: share and rate of the dominant kernel
        plan.profile(True)
        m.logLike(thetas[0], grad=True, exact_grad=True)      # rank-local: no collective outside the timed region
        torch.cuda.synchronize()
        pms, pcnt = plan.profile(False)
        gemm_ms, gemm_launches = float(pms[0].sum()), int(pcnt[0].sum())
        total_ms = float(pms.sum())
        # posterior factorisation (sum cc^2) + its Takahashi pass (2 sum cc^2); the space-time prior is
        # collapsed to two 2-D factorisations (base.py:_prior_collapsed) whose flops are negligible and not counted
        collapsed = bool(getattr(m, "timed", False) and getattr(m, "collapse_prior", False))
        alg_flops = (3.0 if collapsed else 6.0) * stats["flops"]
        peak = fp64_peak()
        achieved = alg_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else None
        traffic, traffic_note = None, "no ncu capture for this workload"
        tpath = os.path.join(ROOT, "profiles", "r1_gemm_traffic_%s.json" % ("c3" if name == "c5" else name))
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            traffic = tj["traffic_bytes_per_launch"]
            traffic_note = ("dram__bytes_read.sum + dram__bytes_write.sum summed over the %d k_gemm_grouped launches of one "
                            "evaluation (%.0f GB read, %.0f GB written) / launches, from %s; the largest launch alone: DMMA pipe "
                            "83%% active, L2 hit 77%% (profiles/r1_ncu_gemm_full_summary.json)"
                            % (tj["launches"], tj["dram_read_bytes"] / 1e9, tj["dram_write_bytes"] / 1e9, os.path.basename(tpath)))
        roof = {"bound": "tensor", "kernel": "k_gemm_grouped (FP64 DMMA m8n8k4)", "achieved": achieved, "peak": peak,
                "unit": "TFLOP/s", "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                "traffic_note": traffic_note,
                "peak_source": "cuBLAS DGEMM 8192^3 measured in this run (MEASURED_PEAKS.json has no FP64 figure)",
                "launches_per_step": gemm_launches, "avg_launch_ms": gemm_ms / max(gemm_launches, 1),
                "gemm_share_of_scheduled_time": gemm_ms / total_ms if total_ms else None,
                "algorithmic_flops_per_step": alg_flops,
                "by_kind_ms": {k: float(pms[i].sum()) for i, k in enumerate(
                    ["gemm", "potrf", "extend_add", "memset", "gather", "wtw", "extract", "gemv"])}}
        # Cholesky GFLOP/s = sum_j cc_j^2 / t_factor (BASELINE.json metric, CHOLMOD's flop convention)
        Qdev = m._state["Q"]
        m.engine.factorize(1, Qdev)
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        f0.record()
        for _ in range(3):
            m.engine.factorize(1, Qdev)
        f1.record()
        torch.cuda.synchronize()
        chol_gflops = stats["flops"] / (f0.elapsed_time(f1) / 3 * 1e-3) / 1e9
        cpu = cpu_reference(name, stats) if world == 1 and not args.no_cpu else None
        value = world * args.steps / (ms * 1e-3)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOADS[name][8], "mesh": [inp["M"], inp["N"], inp["T"]], "n": inp["n"],
                       "npar": int(npar), "nobs": int(inp["idx"].size), "gradient": "exact (Takahashi selected inversion)",
                       "parallelism": "theta-parallel x%d, replicated factorisations" % world,
                       "l2": "inputs larger than L2 (factor %.1f GB)" % (stats["factor_bytes"] / 1e9)},
            "cholesky_gflops": chol_gflops,
            "symbolic": {k: stats[k] for k in ("nsuper", "nnzL", "flops", "factor_bytes", "levels", "max_front")},
            "e2e": {"value": world * args.steps / (ms_e2e * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": cnt_e2e["h2d"] // args.steps, "d2h_bytes_per_step": cnt_e2e["d2h"] // args.steps},
            "gpu_launches": launches, "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
            "last_like": float(last[0]),
        }
        if name == "c3":
            # the mesh BASELINE.json quotes its metric on needs minutes per evaluation (streamed), so it
            # is a separate command and not this default line
            line["config"]["headline_mesh"] = ("256x256x100 (configs[3]) runs on one B200 through the streamed evaluator: "
                                               "python bench.py --workload c4 --steps 1 --warmup 1; measured lines under "
                                               "profiles/ (r1_bench_c4_overlap.json)")
        if nsamp:
            line["config"]["samples_per_theta"] = nsamp
            line["mean_sample_variance"] = sample_chk[0]
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return line


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from spdepy_b200 import _lib
    name = args.workload
    spde, spde0, ha, ani, bc, M, N, T, desc = WORKLOADS[name]
    stats = _lib.PlanHandle(M, N, T or 1, bc).stats()
    vals = []
    for _ in range(max(1, min(args.steps, 3))):
        vals.append(cpu_reference(name, stats))
    best = max(vals, key=lambda c: c["value"])
    inp = make_inputs(name, 4, 4, 2 if T else None)
    line = {
        "impl": "reference", "metric": METRIC, "value": best["value"], "unit": UNIT, "n_gpus": int(os.environ.get("WORLD_SIZE", "1")),
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / best["value"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "mesh": [M, N, T], "n": M * N * (T or 1), "npar": int(inp["theta"].size),
                   "gradient": "Hutchinson nh1=100 (the reference's estimator)"},
        "cpu_baseline": best,
        "e2e": {"value": best["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))   # c5 = batched sweep (configs[4])
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--streamed", action="store_true", help="force the streamed (depth-first) evaluator on any workload")
    ap.add_argument("--mesh", type=int, nargs=3, default=None, help="override the mesh of the workload (streamed runs)")
    ap.add_argument("--profile-step", action="store_true", help="streamed runs: one extra evaluation with per-launch events")
    ap.add_argument("--check-forward", action="store_true", help="streamed runs: also evaluate logLike(grad=False) (forward pass only)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
