"""bench.py -- log-likelihood + gradient evaluations per second of the SPDE hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c2|c1|c4|c5] [--impl ours|reference]

Metric (BASELINE.json): "loglik+grad evals/sec and Cholesky GFLOP/s at 256x256x100 ... vs host-core reference".  One
*step* is one ``logLike(theta, grad=True, exact_grad=True)`` of the workload's model: assemble Q, factorise
Q + tau S^T S (3-D) and the two 2-D matrices the prior's determinant collapses to, logdets, conditional mean, quadratic
forms, Takahashi selected inverses, gradient contraction.

What the default line (``--gpus 1``) holds:

* the timed region of ``--steps K --warmup W`` on configs[2] of BASELINE.json (var-advection-var-diffusion on the
  SINMOD-shaped 100x100x50 mesh, n = 5e5, 92 parameters): ``value``, ``e2e``, ``roofline``, ``clocks`` -- the largest
  config whose evaluation is short enough for K = 20 steps;
* a ``c5`` block: BASELINE configs[4], a FIXED batch of 64 theta x 1024 samples on the same mesh, sharded over the ranks
  (strong scaling: the block is measured at every N, so the per-N lines give the batched scaling);
* a ``c4`` block (N = 1 only): BASELINE configs[3] itself -- ONE evaluation of logLike + exact gradient on the
  256x256x100 mesh (the mesh the metric is quoted on) through the streamed evaluator after one warm-up evaluation.  It
  runs on one B200 but takes about two minutes per evaluation, so it cannot be the K-step region;
  it is inside the same process, so the driver's wall clock bounds it.  Flat copies of its numbers are in
  ``roofline.c4_*`` (and of the c5 block in ``e2e.c5_*``).

N > 1 (torchrun, one rank per GPU): the K-step region is weak scaling over independent theta evaluations, each rank
owning its own factorisation; the only collective is one NCCL all-reduce of the npar+1 likelihood / gradient scalars per
step (BASELINE.json north_star).

``--impl reference``: the reference's CPU path, MEASURED: one real ``logLike(theta, nh1=100, grad=True)`` of the oracle
port (oracle/spde_oracle.py: the reference's algorithm in SciPy, every dQ_i materialised, Hutchinson probes;
oracle/cpu_cholesky.py + oracle/symbolic_oracle.py: supernodal Cholesky on LAPACK standing in for the absent CHOLMOD)
on the FULL workload with all host threads.  It takes minutes, so ``steps`` is reported as the number of evaluations
actually run (1), never the K requested.  This arm imports nothing from ``spdepy_b200``.
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
REF_CACHE = "/tmp/spde_b200_reference_arm_%s.json"

WORKLOADS = {
    # name: (spde, mod0 spde, ha, ani, bc, M, N, T, description)
    "c1": ("whittle-matern", None, False, True, 3, 30, 30, None, "stationary Whittle-Matern 30x30 (configs[0])"),
    "c2": ("advection-diffusion", "whittle-matern", False, True, 3, 50, 50, 20,
           "advection-diffusion 50x50x20, 5000 obs (configs[1])"),
    "c3": ("var-advection-var-diffusion", "var-whittle-matern", False, True, 1, 100, 100, 50,
           "var-advection-var-diffusion 100x100x50 SINMOD-shaped, 92 parameters, 10% obs (configs[2])"),
    # batched sweep: a fixed batch of 64 thetas, each one theta-evaluation (likelihood + exact gradient) PLUS 1024 prior
    # samples at that theta (3-D factor of Q, 1024-column back substitution); thetas are sharded over the ranks
    "c5": ("var-advection-var-diffusion", "var-whittle-matern", False, True, 1, 100, 100, 50,
           "batched sweep on 100x100x50: 64 thetas x (logLike+exact gradient and 1024 samples), thetas sharded over the GPUs (configs[4])"),
    # half-resolution version of configs[3]: same model, cell size, theta and 1 % observation density on 128x128x50
    "c4h": ("advection-diffusion", "whittle-matern", False, True, 3, 128, 128, 50,
            "advection-diffusion 128x128x50, 1% obs: half-resolution version of configs[3]"),
    # configs[3] itself, the mesh the metric is quoted on: one B200, streamed evaluation (spdepy_b200/csrc/ooc.cu) -- the
    # 260 GB factor never lives in HBM as a whole.  ~2 minutes per evaluation: as a workload of its own it is run with
    # `--workload c4 --steps 1 --warmup 1`; the default line carries it as the `c4` block.
    "c4": ("advection-diffusion", "whittle-matern", False, True, 3, 256, 256, 100,
           "advection-diffusion 256x256x100, 1% obs (configs[3]), one B200, streamed depth-first evaluation with pinned-host panel store"),
}
C5_THETAS, C5_SAMPLES = 64, 1024


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


def config_of(name, inp, world, gradient="exact"):
    """The `config` object -- identical in both arms (`--impl ours` / `--impl reference`) for the same command line."""
    grad = {"exact": "logLike(theta, grad=True): exact traces by Takahashi selected inversion on the GPU arm; the reference arm "
                     "runs the reference's own estimator (Hutchinson, nh1=100)",
            "hutchinson": "logLike(theta, nh1=100, grad=True): the reference's default Hutchinson estimator in both arms"}[gradient]
    return {"workload": WORKLOADS[name][8], "mesh": "%dx%dx%s" % (inp["M"], inp["N"], inp["T"] or 1), "n": int(inp["n"]),
            "npar": int(inp["theta"].size), "nobs": int(inp["idx"].size), "gradient": grad,
            "parallelism": "theta-parallel x%d, replicated factorisations" % world,
            "l2": "inputs larger than L2 (126 MB): the factor alone is GBs; every step assembles and factorises afresh"}


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


# ---------------------------------------------------------------------------------------------
# the reference's CPU path: the oracle port, nothing from spdepy_b200

def build_oracle(inp):
    import spde_oracle as so
    from grid_oracle import OracleGrid
    g = OracleGrid(inp["x"], inp["y"], inp["t"])
    names = {"whittle-matern": "whittle-matern-%s-2D", "var-whittle-matern": "var-whittle-matern-%s-2D"}
    if inp["t"] is None:
        return so.OracleSPDE(names[inp["spde"]] % "anisotropic", g, bc=inp["bc"]), g
    g0 = OracleGrid(inp["x"], inp["y"])
    o0 = so.OracleSPDE(names[inp["spde0"]] % ("isotropic" if inp["iso0"] else "anisotropic"), g0, bc=inp["bc"], par=inp["p0"])
    return so.OracleSPDE(inp["spde"] + "-2D", g, mod0=o0, bc=inp["bc"]), g


class OracleFactoriser:
    """``cholesky(A)`` of the port: nested dissection + symbolic analysis once per matrix size (the reference's CHOLMOD
    analyses on EVERY call, advection_diffusion2D.py:117,193 -- reuse is in the CPU's favour), numeric supernodal
    factorisation on LAPACK.  Records the wall time of each phase."""

    def __init__(self, shapes):
        self.shapes = shapes          # n -> (M, N, T, bc)
        self.sym, self.t_symbolic, self.t_numeric, self.nfact, self.sizes = {}, 0.0, 0.0, 0, []

    def __call__(self, A, perm=None):
        import cpu_cholesky as cc
        import symbolic_oracle as syo
        n = A.shape[0]
        if n not in self.sym:
            t0 = time.perf_counter()
            M, N, T, bc = self.shapes[n]
            self.sym[n] = syo.OracleSymbolic(A, syo.nd_perm(M, N, T, bc))
            self.t_symbolic += time.perf_counter() - t0
        t0 = time.perf_counter()
        f = cc.SupernodalFactor(A, plan=self.sym[n])
        self.t_numeric += time.perf_counter() - t0
        self.nfact += 1
        self.sizes.append(n)
        return f


def host_threads():
    """Every host core for the dense kernels of the port, regardless of OMP_NUM_THREADS (torchrun exports
    OMP_NUM_THREADS=1).  The port's dense work goes through scipy.linalg (its own OpenBLAS); numpy's separate OpenBLAS pool
    is parked at one thread so that two pools of spinning workers do not fight for the cores."""
    n = os.cpu_count() or 1
    try:
        import scipy.linalg  # noqa: F401  (loads scipy's BLAS so that it is listed)
        from threadpoolctl import ThreadpoolController
        for lc in ThreadpoolController().lib_controllers:
            lc.set_num_threads(n if "scipy.libs" in (lc.filepath or "") else 1)
    except Exception:
        pass
    return n


def reference_eval(name, mesh=None, nh1=100):
    """ONE real evaluation of the port: ``logLike(theta, nh1, grad=True)`` as advection_diffusion2D.py:187-223 does it
    (assembly of Q and of every dQ_i in SciPy, cholesky(Q), cholesky(Q_c), 2 x nh1-column solves, npar SpMMs).
    Returns the wall time and a breakdown; ``mesh`` (M, N, T) shrinks the workload for the bounded cpu_baseline sample."""
    import spde_oracle as so
    cores = host_threads()
    inp = make_inputs(name, *mesh) if mesh else make_inputs(name)
    bc, Ns = inp["bc"], inp["M"] * inp["N"]
    fac = OracleFactoriser({inp["n"]: (inp["M"], inp["N"], inp["T"] or 1, bc), Ns: (inp["M"], inp["N"], 1, bc)})
    so.set_factor(fac)
    try:
        t0 = time.perf_counter()
        mod, g = build_oracle(inp)      # (builds and factorises the initial-field model mod0 once, as sp.model() does)
        mod.initFit(inp["data"], idx=inp["idx"])
        t_build = time.perf_counter() - t0
        fac.t_symbolic_build, fac.t_numeric_build, fac.nfact_build = fac.t_symbolic, fac.t_numeric, fac.nfact
        nsizes_build = len(fac.sizes)
        np.random.seed(4)
        t0 = time.perf_counter()
        like, jac = mod.logLike(inp["theta"], nh1=nh1, grad=True)
        t_eval = time.perf_counter() - t0
    finally:
        so.set_factor(None, None)
    st = fac.sym[g.n].stats()
    t_sym, t_num, nf = fac.t_symbolic - fac.t_symbolic_build, fac.t_numeric - fac.t_numeric_build, fac.nfact - fac.nfact_build
    return {"seconds": t_eval, "like": float(like), "jac_inf": float(np.abs(jac).max()), "cores": cores, "inp": inp,
            "t_build_model": t_build, "t_symbolic": t_sym, "t_numeric_factor": t_num, "factorisations": nf,
            "stats": st, "cpu_cholesky_gflops": st["flops"] * 2 / max(t_num, 1e-9) / 1e9,
            "full_size_factorisations": sum(1 for v in fac.sizes[nsizes_build:] if v == g.n)}


def cpu_baseline(name):
    """cpu_baseline of the GPU arm.  Preferred: the measurement `bench.py --impl reference` made on this very box (the
    driver runs that arm first); otherwise a bounded sample: one real evaluation of the port on a reduced mesh of the same
    model, scaled to the workload (assembly + SpMMs by n, factorisations by sum cc^2, the rest by nnz(L))."""
    path = REF_CACHE % name
    try:
        if os.path.exists(path) and time.time() - os.path.getmtime(path) < 6 * 3600:
            c = json.load(open(path))
            c["sample"] = "full workload, measured by `bench.py --impl reference` on this box %.0f s earlier: %s" % (
                time.time() - os.path.getmtime(path), c["sample"])
            return c
    except Exception:
        pass
    spde, spde0, ha, ani, bc, M0, N0, T0, _ = WORKLOADS[name]
    if T0 is None or M0 * N0 * T0 <= 100000:      # small enough to run in full (c1: 0.03 s, c2: ~10 s)
        r = reference_eval(name)
        return {"value": 1.0 / r["seconds"], "unit": UNIT, "cores": r["cores"], "kind": "port", "scaled": False,
                "sample": "full workload: one logLike(grad=True, nh1=100) of the oracle port, %.2f s" % r["seconds"]}
    mesh = (40, 40, 16)
    r = reference_eval(name, mesh)
    # full-size symbolic quantities from the oracle's own analysis (pattern only)
    t0 = time.perf_counter()
    full = _oracle_stats(M0, N0, T0, bc)
    t_sym = time.perf_counter() - t0
    s = r["stats"]
    t_fac = r["t_numeric_factor"]
    t_rest = r["seconds"] - t_fac - r["t_symbolic"]
    # the small factorisations of the sample run far below the host's dense rate, so scaling THEIR time by flops would
    # overstate the CPU tenfold: the full-size factorisations are charged at the dense LAPACK Cholesky rate measured here
    # (n^3/3 = sum cc^2 of a dense matrix; optimistic for the CPU: the measured supernodal rate at full size is ~70 % of it)
    rate = _host_potrf_rate()
    nbig = max(r.get("full_size_factorisations", 2), 1)
    t_fac_full = nbig * full["flops"] / rate
    full_t = t_fac_full + t_rest * full["n"] / s["n"] + t_sym
    return {"value": 1.0 / full_t, "unit": UNIT, "cores": r["cores"], "kind": "port", "scaled": True,
            "sample": ("bounded sample (no reference-arm measurement found on this box): one real logLike(grad=True, nh1=100) of the "
                       "oracle port on a %dx%dx%d mesh of the same model took %.1f s (%.1f s in %d numeric factorisations); scaled to "
                       "the workload: %d full-size factorisations of sum cc^2 = %.3g at the dense LAPACK Cholesky rate of this host "
                       "(%.0f GFLOP/s, n^3/3 convention) = %.0f s, everything else by n (x%.0f) = %.0f s, symbolic %.1f s -> %.0f s per "
                       "evaluation.  The measured full-size figure is the `--impl reference` line (164-168 s on this pool's boxes)" %
                       (mesh + (r["seconds"], t_fac, r["factorisations"], nbig, full["flops"], rate / 1e9, t_fac_full, full["n"] / s["n"],
                                t_rest * full["n"] / s["n"], t_sym, full_t))),
            "sample_seconds": r["seconds"]}


def _host_potrf_rate(n=5000):
    """Dense Cholesky rate of this host, flops = n^3/3 (the sum cc^2 of a dense matrix), best of two."""
    import scipy.linalg as sla
    rng = np.random.default_rng(0)
    A = rng.normal(size=(n, n))
    A = A @ A.T + n * np.eye(n)
    best = 1e30
    for _ in range(2):
        t0 = time.perf_counter()
        sla.cholesky(A, lower=True, overwrite_a=False, check_finite=False)
        best = min(best, time.perf_counter() - t0)
    return n ** 3 / 3.0 / best


def _oracle_stats(M, N, T, bc):
    """nnz(L), sum cc^2 of the mesh pattern from the oracle's symbolic analysis (no numeric work)."""
    import symbolic_oracle as syo
    from scipy import sparse
    Ns, n = M * N, M * N * T
    k = np.arange(n)
    x, y, t = k % M, (k // M) % N, k // Ns
    rows, cols = [], []
    for dtt, rad in ((0, 2), (1, 1)):
        for dy in range(-rad, rad + 1):
            for dx in range(-rad, rad + 1):
                xx, yy, tt = x + dx, y + dy, t + dtt
                if bc == 2:
                    xx, yy = xx % M, yy % N
                ok = (xx >= 0) & (xx < M) & (yy >= 0) & (yy < N) & (tt < T)
                rows.append(k[ok]); cols.append((tt * Ns + yy * M + xx)[ok])
    r, c = np.concatenate(rows), np.concatenate(cols)
    A = sparse.csc_matrix((np.ones(r.size, dtype=np.int8), (r, c)), shape=(n, n))
    return syo.OracleSymbolic(A, syo.nd_perm(M, N, T, bc)).stats()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    name = args.workload
    t_wall = time.time()
    r = reference_eval(name, tuple(args.mesh) if args.mesh else None)
    inp = r["inp"]
    st = r["stats"]
    value = 1.0 / r["seconds"]
    cb = {"value": value, "unit": UNIT, "cores": r["cores"], "kind": "port",
          "sample": ("the FULL workload, once: logLike(theta, nh1=100, grad=True) of the oracle port (SciPy assembly of Q and of all "
                     "%d dQ_i as the reference does, supernodal Cholesky on LAPACK with %d threads standing in for CHOLMOD, two "
                     "100-column solves, npar SpMMs): %.1f s, of which symbolic %.1f s, %d numeric factorisations %.1f s "
                     "(%.0f GFLOP/s); nnz(L) %.3g, sum cc^2 %.3g"
                     % (inp["theta"].size - 1, r["cores"], r["seconds"], r["t_symbolic"], r["factorisations"], r["t_numeric_factor"],
                        r["cpu_cholesky_gflops"], st["nnzL"], st["flops"])),
          "seconds_per_eval": r["seconds"], "like": r["like"]}
    if not args.mesh:
        try:
            json.dump(cb, open(REF_CACHE % name, "w"))
        except Exception:
            pass
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
        "steps": 1, "warmup": 0, "steps_requested": args.steps, "warmup_requested": args.warmup,
        "ms_per_step": 1e3 * r["seconds"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_of(name, inp, world, args.gradient),
        "cpu_baseline": cb,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "one evaluation takes minutes on the host, so exactly ONE was run and timed (steps = 1, warmup = 0) whatever "
                "--steps/--warmup asked for; rank 0 only under torchrun; wall time of this arm %.0f s" % (time.time() - t_wall),
    }
    print(json.dumps(line))


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


_FP64_PEAK = [None]


def fp64_peak():
    """cuBLAS DGEMM 8192^3 through torch.matmul: the FP64 roofline denominator (MEASURED_PEAKS.json has
    HBM and bf16 only).  Burst figure, best of 3."""
    if _FP64_PEAK[0] is not None:
        return _FP64_PEAK[0]
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
    _FP64_PEAK[0] = 2 * n ** 3 / best / 1e9     # TFLOP/s
    return _FP64_PEAK[0]


def streamed_block(args, name="c4", steps=1, warmup=1, mesh=None, profile_step=False, check_forward=False):
    """configs[3] (256x256x100) on one B200: `steps` evaluations of logLike + exact gradient with the streamed evaluator,
    after `warmup` evaluations.  Returns a dict of measurements; `roofline_achieved` is the whole-pass rate (algorithmic
    flops of the posterior factorisation + Takahashi pass / device time of the two passes), a lower bound of the GEMM rate."""
    import torch
    from spdepy_b200 import _lib
    from spdepy_b200.engine import COUNTERS
    t0 = time.time()
    inp = make_inputs(name, *mesh) if mesh else make_inputs(name)
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
    print("# %s: plan %.1f s, streamed plan %.1f s: %s" % (name, t_plan, t_ooc, json.dumps(ost)), file=sys.stderr, flush=True)
    theta = inp["theta"]
    passes = []

    def step():
        like, jac = m.logLike(theta, grad=True, exact_grad=True)
        passes.append((ooc.info_d(1), ooc.info_d(2)))
        return like, jac

    t0 = time.time()
    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    t_warm = time.time() - t0
    COUNTERS["h2d"] = COUNTERS["d2h"] = 0
    _lib.lib.spde_launch_count(1)
    with ClockSampler(torch.cuda.current_device()) as clk:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            like, jac = step()
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    launches = int(_lib.lib.spde_launch_count(0))
    h2d, d2h = COUNTERS["h2d"] // max(steps, 1), COUNTERS["d2h"] // max(steps, 1)
    fwd_ms, bwd_ms = passes[-1]
    print("# %s evaluation: %.1f s (forward %.1f s, backward %.1f s), like %.12g" % (name, ms / steps / 1e3, fwd_ms / 1e3, bwd_ms / 1e3, like),
          file=sys.stderr, flush=True)
    # full-size checks: residual of the conditional mean, Q_c mu = tau S^T y, and tr(Q_c Z)/n of the selected inverse
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
    if profile_step:
        # one more evaluation with per-launch CUDA events (outside the timed region): device time per launch kind
        plan.profile(True)
        m.logLike(theta, grad=True, exact_grad=True)
        torch.cuda.synchronize()
        pms, pcnt = plan.profile(False)
        kinds = ["gemm", "potrf", "extend_add", "memset", "gather", "wtw", "extract", "gemv"]
        by_kind = {"ms": {k: float(pms[i].sum()) for i, k in enumerate(kinds)},
                   "launches": {k: int(pcnt[i].sum()) for i, k in enumerate(kinds)},
                   "passes_ms": [ooc.info_d(1), ooc.info_d(2)],
                   "gemm_tflops": (3.0 * stats["flops"] + ost["recompute_flops"]) / (float(pms[0].sum()) * 1e-3) / 1e12,
                   "note": "profiled evaluation (events around every launch); gemm_tflops = (3 sum cc^2 + flops factorised twice) / "
                           "summed k_gemm_grouped time; host<->device panel copies and update-matrix moves are not launches: "
                           "their time is passes_ms minus the sum over kinds"}
    like_fwd = float(m.logLike(theta, grad=False)) if check_forward else None
    alg = 3.0 * stats["flops"]
    peak = fp64_peak()
    achieved = alg / ((fwd_ms + bwd_ms) * 1e-3) / 1e12
    out = {
        "workload": WORKLOADS[name][8], "mesh": "%dx%dx%d" % (inp["M"], inp["N"], inp["T"]), "n": int(inp["n"]),
        "npar": int(theta.size), "nobs": int(inp["idx"].size), "steps": steps, "warmup": warmup,
        "evals_per_s": steps / (ms * 1e-3), "ms_per_eval": ms / steps,
        "cholesky_gflops": stats["flops"] / (fwd_ms * 1e-3) / 1e9, "forward_pass_ms": fwd_ms, "backward_pass_ms": bwd_ms,
        "roofline_achieved_tflops": achieved, "roofline_peak_tflops": peak, "roofline_frac": achieved / peak,
        "algorithmic_flops_per_eval": alg,
        "conditional_mean_residual_inf": resid, "selinv_trace_over_n": trace_check, "like": float(like),
        "like_forward_only": like_fwd, "grad_inf_norm": float(np.abs(jac).max()),
        "clocks": clk.summary(), "h2d_bytes_per_eval": h2d, "d2h_bytes_per_eval": d2h, "gpu_launches": launches,
        "host_setup_s": {"symbolic_and_plan": t_plan, "streamed_plan": t_ooc, "warmup_evaluations": t_warm},
        "segments": ost["segments"], "top_segments": ost["top_segments"], "device_pool_gb": ost["pool_bytes"] / 1e9,
        "pinned_host_gb": ost["host_bytes"] / 1e9, "recompute_flops": ost["recompute_flops"],
        "symbolic": {k: stats[k] for k in ("nsuper", "nnzL", "flops", "factor_bytes", "levels", "max_front")},
        "streamed": ost, "profile_by_kind": by_kind, "passes_ms_per_eval": [list(p) for p in passes],
        "jac": [float(v) for v in jac],
        "note": "whole-pass roofline: 3 sum cc^2 / (forward + backward device time), includes scatter, extend-add, host transfers "
                "and the recomputed subtrees -- a lower bound of k_gemm_grouped's own rate; the pinned-host panel traffic "
                "(%.1f GB each way per evaluation) is inside the timed region" % (ost["host_bytes"] / 1e9),
    }
    # release the pool (most of the device) before anything else runs
    del mod, m, ooc, Q, mu
    eng._ooc = None
    from spdepy_b200.engine import Engine
    Engine._cache.clear()
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    return out


def run_streamed(args):
    """`--workload c4` (or `--streamed`): the streamed evaluator as a workload of its own; same JSON line as the others."""
    import torch
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    name = args.workload
    blk = streamed_block(args, name, steps=args.steps, warmup=args.warmup, mesh=tuple(args.mesh) if args.mesh else None,
                         profile_step=args.profile_step, check_forward=args.check_forward)
    inp = make_inputs(name, *args.mesh) if args.mesh else make_inputs(name)
    cfg = config_of(name, inp, 1)
    cfg["parallelism"] = "one GPU, streamed: %d segments (%d front-by-front), device pool %.1f GB, pinned host %.1f GB" % (
        blk["segments"], blk["top_segments"], blk["device_pool_gb"], blk["pinned_host_gb"])
    line = {
        "metric": METRIC, "value": blk["evals_per_s"], "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": blk["ms_per_eval"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": cfg, "cholesky_gflops": blk["cholesky_gflops"],
        "e2e": {"value": blk["evals_per_s"], "unit": UNIT, "h2d_bytes_per_step": blk["h2d_bytes_per_eval"],
                "d2h_bytes_per_step": blk["d2h_bytes_per_eval"]},
        "gpu_launches": blk["gpu_launches"], "clocks": blk["clocks"],
        "roofline": {"bound": "tensor", "kernel": "k_gemm_ws + k_gemm_grouped (FP64 DMMA m8n8k4; bulk-async warp-specialised tiles / cp.async ring tiles)", "achieved": blk["roofline_achieved_tflops"],
                     "peak": blk["roofline_peak_tflops"], "unit": "TFLOP/s", "frac": blk["roofline_frac"], "traffic": None,
                     "peak_source": "cuBLAS DGEMM 8192^3 measured in this run", "note": blk["note"]},
        "cpu_baseline": None if args.no_cpu else cpu_baseline(name),
        "c4": blk,
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
    name = args.workload
    hutch = args.gradient == "hutchinson"
    line, c5 = incore_region(args, rank, world, local)
    # ---- c4 block (N = 1): BASELINE configs[3], the mesh the metric is quoted on, one evaluation after one warm-up
    c4 = None
    if world == 1 and name == "c3" and not args.no_c4 and not hutch:
        from spdepy_b200.engine import Engine
        import gc
        Engine._cache.clear()
        gc.collect()
        torch.cuda.empty_cache()
        try:
            c4 = streamed_block(args, "c4", steps=1, warmup=1)
        except Exception as e:      # the default line must still be printed
            c4 = {"error": "%s: %s" % (type(e).__name__, e)}
        flat = {"c4_" + k: v for k, v in c4.items() if not isinstance(v, (dict, list)) and k not in ("note", "workload")}
        if "clocks" in c4:
            flat["c4_sm_mhz"] = c4["clocks"]["sm_mhz"]
            flat["c4_clock_reasons"] = ",".join(c4["clocks"]["reasons"])
        line["roofline"].update(flat)
    if rank == 0:
        line["cpu_baseline"] = cpu_baseline(name) if world == 1 and not args.no_cpu else None
        if c5 is not None:
            line["c5"] = c5
        if c4 is not None:
            line["c4"] = {k: v for k, v in c4.items() if k not in ("streamed", "jac", "passes_ms_per_eval")}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return line


def incore_region(args, rank, world, local):
    """The K-step timed region (device-resident and end-to-end), the c5 block and the profiled step of an in-core
    workload.  Returns (line or None on ranks > 0, c5 block or None); every device object dies with this frame."""
    import torch
    import torch.distributed as dist
    from spdepy_b200 import _lib
    from spdepy_b200.engine import COUNTERS, to_dev

    name = args.workload
    hutch = args.gradient == "hutchinson"
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

    def evaluate(theta):
        if hutch:       # the reference's default call: probes from the global legacy RNG on the host (advection_diffusion2D.py:200)
            return m.logLike(theta, nh1=100, grad=True)
        return m.logLike(theta, grad=True, exact_grad=True)

    def step(theta, resident):
        if resident:
            if "data" not in m._obs:
                m._obs["data"] = to_dev(inp["data"])
        else:
            m._obs.pop("data", None)
            m.data = pinned_data.numpy()
        like, jac = evaluate(theta)
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

    # ---- c5 block: a FIXED batch of 64 thetas x 1024 samples on the same mesh, thetas sharded over the ranks (strong scaling)
    c5 = None
    if name in ("c3", "c5") and not args.no_c5 and not hutch:
        c5 = c5_block(args, m, inp, rank, world, local)
    if rank != 0:
        return None, c5

    # one profiled step outside the timed region: share and rate of the dominant kernel
    plan.profile(True)
    evaluate(thetas[0])      # rank-local: no collective outside the timed region
    torch.cuda.synchronize()
    pms, pcnt = plan.profile(False)
    gemm_ms, gemm_launches = float(pms[0].sum()), int(pcnt[0].sum())
    total_ms = float(pms.sum())
    # exact gradient: posterior factorisation (sum cc^2) + its Takahashi pass (2 sum cc^2); the space-time
    # prior is collapsed to two 2-D factorisations (base.py:_prior_collapsed) whose flops are negligible and not counted.
    # Hutchinson: two 3-D factorisations + 2 x 100-column solve_A (4 nnz(L) k flop each)
    collapsed = bool(getattr(m, "timed", False) and getattr(m, "collapse_prior", False))
    if hutch:
        alg_flops = 2.0 * stats["flops"] + 2 * 4.0 * stats["nnzL"] * 100
    else:
        alg_flops = (3.0 if collapsed else 6.0) * stats["flops"]
    peak = fp64_peak()
    achieved = alg_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else None
    traffic, traffic_note = None, "no ncu capture for this workload"
    for tag in ("r2", "r1"):
        tpath = os.path.join(ROOT, "profiles", "%s_gemm_traffic_%s.json" % (tag, "c3" if name == "c5" else name))
        if os.path.exists(tpath) and not hutch:
            tj = json.load(open(tpath))
            traffic = tj["traffic_bytes_per_launch"]
            traffic_note = ("dram__bytes_read.sum + dram__bytes_write.sum summed over the %d dense-tile launches (k_gemm_ws, k_gemm_grouped) of one "
                            "evaluation (%.0f GB read, %.0f GB written) / launches, from %s"
                            % (tj["launches"], tj["dram_read_bytes"] / 1e9, tj["dram_write_bytes"] / 1e9, os.path.basename(tpath)))
            break
    by_kind = {k: float(pms[i].sum()) for i, k in enumerate(
        ["gemm", "potrf", "extend_add", "memset", "gather", "wtw", "extract", "gemv"])}
    roof = {"bound": "tensor", "kernel": "k_gemm_ws + k_gemm_grouped (FP64 DMMA m8n8k4; bulk-async warp-specialised tiles / cp.async ring tiles)", "achieved": achieved, "peak": peak,
            "unit": "TFLOP/s", "frac": (achieved / peak) if achieved else None, "traffic": traffic,
            "traffic_note": traffic_note,
            "peak_source": "cuBLAS DGEMM 8192^3 measured in this run (MEASURED_PEAKS.json has no FP64 figure)",
            "launches_per_step": gemm_launches, "avg_launch_ms": gemm_ms / max(gemm_launches, 1),
            "gemm_share_of_scheduled_time": gemm_ms / total_ms if total_ms else None,
            "algorithmic_flops_per_step": alg_flops, "non_gemm_ms_per_step": total_ms - gemm_ms}
    roof.update({"ms_" + k: v for k, v in by_kind.items()})
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
    factor_ms = f0.elapsed_time(f1) / 3
    chol_gflops = stats["flops"] / (factor_ms * 1e-3) / 1e9
    roof["cholesky_gflops"] = chol_gflops
    roof["factor_ms"] = factor_ms
    value = world * args.steps / (ms * 1e-3)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": config_of(name, inp, world, args.gradient),
        "cholesky_gflops": chol_gflops,
        "symbolic": {k: stats[k] for k in ("nsuper", "nnzL", "flops", "factor_bytes", "levels", "max_front")},
        "e2e": {"value": world * args.steps / (ms_e2e * 1e-3), "unit": UNIT,
                "h2d_bytes_per_step": cnt_e2e["h2d"] // args.steps, "d2h_bytes_per_step": cnt_e2e["d2h"] // args.steps},
        "gpu_launches": launches, "clocks": clocks, "roofline": roof,
        "last_like": float(last[0]),
    }
    if c5 is not None:
        line["e2e"].update({"c5_" + k: v for k, v in c5.items() if not isinstance(v, (dict, list)) and k != "workload"})
    return line, c5


def c5_block(args, m, inp, rank, world, local):
    """BASELINE configs[4]: 64 thetas x (logLike + exact gradient, 3-D factor of the prior, 1024 prior samples by
    back substitution) on the 100x100x50 mesh, the fixed batch sharded over the ranks (theta i -> rank i mod N), one NCCL
    all-reduce of the batch's summed likelihood / gradient scalars.  Device time, max over ranks."""
    import torch
    import torch.distributed as dist
    from spdepy_b200 import _lib
    rng = np.random.default_rng(7)
    batch = inp["theta"] + 0.01 * rng.normal(size=(C5_THETAS, inp["theta"].size))
    mine = list(range(rank, C5_THETAS, world))
    eng = m.engine
    npar = inp["theta"].size
    acc = torch.zeros(npar + 2, dtype=torch.float64, device="cuda")      # like, jac, sum of squares of the samples
    gen = torch.Generator(device="cuda")
    blk = 512

    marks = []        # (phase, event) pairs of the timed thetas: where the batch time goes (read after the final sync)

    def mark(phase):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        marks.append((phase, e))

    def one(theta, seed):
        mark("start")
        like, jac = m.logLike(theta, grad=True, exact_grad=True)
        mark("loglike_grad")
        # Model.sample at this theta (model.py:73-87): x = P^T L^-T z with L the factor of the 3-D prior
        eng.factorize(0, m._state["Q"])
        mark("factor_prior")
        gen.manual_seed(8 + seed)
        for c0 in range(0, C5_SAMPLES, blk):
            z = torch.randn(eng.n, min(blk, C5_SAMPLES - c0), dtype=torch.float64, device="cuda", generator=gen)
            mark("draws")
            x = eng.solve(0, z, 10)
            acc[npar + 1] += torch.linalg.vector_norm(x) ** 2        # device-side checksum, no host sync
            mark("back_substitution")
        acc[0] += like
        acc[1:npar + 1] += torch.as_tensor(jac, device="cuda")

    for w in range(2):                   # warm-up: the first call builds the 512-column solve schedules, the second captures their graphs
        one(inp["theta"], 10 ** 6 + w)
    acc.zero_()
    marks.clear()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    _lib.lib.spde_launch_count(1)
    with ClockSampler(local) as clk:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in mine:
            one(batch[i], i)
        if world > 1:
            dist.all_reduce(acc)
        e1.record()
        torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        tms = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms.item())
    h = acc.cpu().numpy()
    phase_ms = {}
    for (_, e0), (ph, e1) in zip(marks[:-1], marks[1:]):
        if ph != "start":
            phase_ms[ph] = phase_ms.get(ph, 0.0) + e0.elapsed_time(e1)
    phase_ms = {k: v / len(mine) for k, v in phase_ms.items()}
    return {"workload": WORKLOADS["c5"][8], "phase_ms_per_theta_rank0": phase_ms, "thetas": C5_THETAS, "samples_per_theta": C5_SAMPLES, "n_gpus": world,
            "scaling": "strong", "value": C5_THETAS / (ms * 1e-3), "unit": "theta-evaluations/s (each with 1024 samples)",
            "ms_batch": ms, "ms_per_theta_per_gpu": ms / len(mine), "thetas_per_rank": len(mine),
            "sum_like": float(h[0]), "mean_sample_variance": float(h[npar + 1] / (eng.n * C5_SAMPLES * C5_THETAS)),
            "gpu_launches_rank0": int(_lib.lib.spde_launch_count(0)), "sm_mhz": clk.summary()["sm_mhz"],
            "clock_reasons": ",".join(clk.summary()["reasons"])}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--gradient", default="exact", choices=["exact", "hutchinson"],
                    help="hutchinson: the reference's default logLike(par, nh1=100) path (two 3-D factors, 2 x 100-column solves)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-c4", action="store_true", help="skip the 256x256x100 block of the default line")
    ap.add_argument("--no-c5", action="store_true", help="skip the batched-sweep block of the default line")
    ap.add_argument("--streamed", action="store_true", help="force the streamed (depth-first) evaluator on any workload")
    ap.add_argument("--mesh", type=int, nargs=3, default=None, help="override the mesh of the workload (streamed runs, reference arm)")
    ap.add_argument("--profile-step", action="store_true", help="streamed runs: one extra evaluation with per-launch events")
    ap.add_argument("--check-forward", action="store_true", help="streamed runs: also evaluate logLike(grad=False) (forward pass only)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
