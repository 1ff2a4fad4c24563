"""Golden fixtures ON THE BASELINE CONFIGS, from the UNMODIFIED reference (run in the build container only).

    python oracle/make_golden_baseline.py [c1] [c2] [c3]

Inputs and seeds are the ones SURVEY.md section 8d fixes:

* C1 (``tests/golden/baseline/c1_<H>_bc<b>.npz``, 6 cases): Whittle-Matern on the reference's default 30x30 mesh,
  ``(ha, anisotropic)`` in {(F,F), (F,T), (T,T)} x bc in {3, 1}, theta = class defaults; ``mod.setModel();
  X = mod.sample(n=100, seed=0)``; observations ``idx = sort(default_rng(1).choice(900, 450, replace=False))``;
  ``data = X[idx, :1]`` (r = 1) and ``X[idx, :20]`` (r = 20); probes ``np.random.seed(4)``, nh1 = 100.  The factoriser
  behind the reference is the dense LAPACK stand-in (identity permutation).
* C2 (``c2_bc<b>.npz``): advection-diffusion on 50x50x20 (``x = y = linspace(0, 15, 50)``, ``t = linspace(0, 2, 20)``), mod0 =
  isotropic Whittle-Matern ``[-2, -0.5, log 10]``, theta = ``[-1,-1,1,-1,1,-1,0,-2,-0.5,log 1000]``, 5 000 observations
  ``idx = sort(default_rng(2).choice(50000, 5000, False))``, ``data = sample(n=20, seed=3)[idx]``, nh1 = 100 with
  ``np.random.seed(4)``.  n = 50 000 is beyond the dense stand-in, so the ``sksparse`` shim is the oracle's supernodal
  Cholesky (``cpu_cholesky.SupernodalFactor`` on an ``OracleSymbolic`` plan).  Q itself (24 MB) is committed as SHA-256
  digests of its canonical CSC arrays -- ``tests/test_oracle_baseline.py`` checks that the oracle restatement reproduces
  them BIT FOR BIT -- plus a strided sample of entries; ``mu_c`` as a strided sample of rows plus norms.
* C3 (``c3.npz``): var-advection-var-diffusion on the SINMOD-shaped 100x100x50 mesh with the 92 fitted parameters,
  inputs of ``bench.make_inputs("c3")``: ``logLike(grad=False)`` of the reference, the two log-determinants, digests of Q.
"""
from __future__ import annotations

import hashlib
import os
import sys
import time

import numpy as np
from scipy import sparse

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import ref_harness as rh  # noqa: E402
import cpu_cholesky as cc  # noqa: E402
import symbolic_oracle as syo  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "baseline")


def canon(Q):
    Q = sparse.csc_matrix(Q).copy()
    Q.sum_duplicates()
    Q.eliminate_zeros()
    Q.sort_indices()
    return Q


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def q_digests(Q):
    Q = canon(Q)
    return dict(Q_sha_indptr=digest(Q.indptr.astype(np.int64)), Q_sha_indices=digest(Q.indices.astype(np.int64)),
                Q_sha_data=digest(Q.data.astype(np.float64)), Q_nnz=Q.nnz,
                Q_data_sample=Q.data[::997].copy(), Q_abs_sum=np.abs(Q.data).sum())


class SparseShim:
    """``sksparse.cholmod.cholesky`` stand-in for large n: nested dissection + supernodal Cholesky of the oracle."""

    def __init__(self, shapes):
        self.shapes, self.sym = shapes, {}

    def __call__(self, A):
        n = A.shape[0]
        if n not in self.sym:
            self.sym[n] = syo.OracleSymbolic(A, syo.nd_perm(*self.shapes[n]))
        return cc.SupernodalFactor(A, plan=self.sym[n])


def make_c1():
    sp = rh.load_reference()
    x = np.linspace(2 / 3, 40 - 2 / 3, 30)
    for ha, ani, tag in ((False, False, "iso"), (False, True, "ani"), (True, True, "ha")):
        for bc in (3, 1):
            g = sp.grid(x=x, y=x)
            mod = sp.model(grid=g, spde="whittle-matern", ha=ha, anisotropic=ani, bc=bc)
            par = np.array(mod.mod.getPars(), dtype="float64")
            rh.set_permutation(None)
            mod.mod.setQ(par=par)
            Q = canon(mod.mod.Q)
            mod.setModel()
            X = mod.sample(n=100, seed=0)
            idx = np.sort(np.random.default_rng(1).choice(900, 450, replace=False))
            out = dict(ha=ha, ani=ani, bc=bc, x=x, par=par, type=mod.mod.type, idx=idx, sample20=X[:, :20],
                       sample_colnorms=np.sqrt((X ** 2).sum(axis=0)),
                       Q_data=Q.data, Q_indices=Q.indices.astype(np.int32), Q_indptr=Q.indptr.astype(np.int32))
            from sksparse.cholmod import cholesky
            for r in (1, 20):
                data = X[idx, :r]
                mod.mod.initFit(data, idx=idx)
                like, jac = rh.loglike_seeded(mod.mod, par, nh1=100, grad=True, seed=4)
                tau = np.exp(par[-1])
                Qm, Qf, _ = mod.mod.makeQ(par=par, grad=False)
                S = mod.mod.S
                Qcf = cholesky(Qm + S.T @ S * tau)
                mu_c = Qcf.solve_A(S.T @ data * tau)
                out.update({"like_r%d" % r: like, "jac_r%d" % r: jac, "mu_c_r%d" % r: mu_c.reshape(900, r),
                            "logdetQ": Qf.logdet(), "logdetQc_r%d" % r: Qcf.logdet()})
            name = "c1_%s_bc%d" % (tag, bc)
            np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
            print(name, Q.nnz, out["like_r1"], out["like_r20"])


def make_c2():
    sp = rh.load_reference()
    x = np.linspace(0, 15, 50)
    t = np.linspace(0, 2, 20)
    theta = np.array([-1, -1, 1, -1, 1, -1, 0, -2, -0.5, np.log(1000.0)], dtype="float64")
    for bc in (3, 1):
        rh.set_sparse_factor(SparseShim({50000: (50, 50, 20, bc), 2500: (50, 50, 1, bc)}))
        try:
            g = sp.grid(x=x, y=x, t=t)
            m0 = sp.model(grid=sp.grid(x=x, y=x), spde="whittle-matern", ha=False, anisotropic=False, bc=bc,
                          parameters=np.array([-2.0, -0.5, np.log(10.0)]))
            mod = sp.model(grid=g, spde="advection-diffusion", ha=False, anisotropic=True, bc=bc, mod0=m0)
            t0 = time.time()
            mod.mod.setQ(par=theta)
            Q = canon(mod.mod.Q)
            mod.setModel()
            X = mod.sample(n=20, seed=3)
            idx = np.sort(np.random.default_rng(2).choice(50000, 5000, replace=False))
            data = X[idx]
            mod.mod.initFit(data, idx=idx)
            like, jac = rh.loglike_seeded(mod.mod, theta, nh1=100, grad=True, seed=4)
            like0 = rh.loglike_seeded(mod.mod, theta, nh1=100, grad=False, seed=4)
            from sksparse.cholmod import cholesky
            tau = np.exp(theta[-1])
            Qm, Qf, _ = mod.mod.makeQ(par=theta, grad=False)
            S = mod.mod.S
            Qcf = cholesky(Qm + S.T @ S * tau)
            mu_c = Qcf.solve_A(S.T @ data * tau)
            out = dict(bc=bc, x=x, t=t, par=theta, type=mod.mod.type, idx=idx, data=data, like=like, like_nograd=like0, jac=jac,
                       logdetQ=Qf.logdet(), logdetQc=Qcf.logdet(), mu_c_rows=mu_c[::50].copy(), mu_c_colnorms=np.sqrt((mu_c ** 2).sum(axis=0)),
                       mu_c_sum=mu_c.sum(), **q_digests(Q))
            np.savez_compressed(os.path.join(OUT, "c2_bc%d.npz" % bc), **out)
            print("c2 bc", bc, Q.nnz, like, "%.0f s" % (time.time() - t0))
        finally:
            rh.set_sparse_factor(None)


def make_c3():
    import bench
    sp = rh.load_reference()
    inp = bench.make_inputs("c3")
    bc = inp["bc"]
    rh.set_sparse_factor(SparseShim({inp["n"]: (100, 100, 50, bc)}))
    try:
        t0 = time.time()
        g = sp.grid(x=inp["x"], y=inp["y"], t=inp["t"])
        m0 = sp.model(grid=sp.grid(x=inp["x"], y=inp["y"]), spde="var-whittle-matern", ha=False, anisotropic=True, bc=bc,
                      parameters=inp["p0"])
        mod = sp.model(grid=g, spde="var-advection-var-diffusion", ha=False, anisotropic=True, bc=bc, mod0=m0)
        mod.mod.initFit(inp["data"], idx=inp["idx"], fitQ0=False)
        like = mod.mod.logLike(inp["theta"], grad=False)
        print("c3 logLike(grad=False)", like, "%.0f s" % (time.time() - t0), flush=True)
        from sksparse.cholmod import cholesky
        tau = np.exp(inp["theta"][-1])
        Qm, Qf, _ = mod.mod.makeQ(par=inp["theta"], grad=False)
        S = mod.mod.S
        Qcf = cholesky(Qm + S.T @ S * tau)
        mu_c = Qcf.solve_A(S.T @ inp["data"] * tau)
        out = dict(bc=bc, par=inp["theta"], type=mod.mod.type, like=like, logdetQ=Qf.logdet(), logdetQc=Qcf.logdet(),
                   mu_c_rows=mu_c.reshape(-1)[::500].copy(), mu_c_norm=np.sqrt((mu_c ** 2).sum()), **q_digests(Qm))
        np.savez_compressed(os.path.join(OUT, "c3.npz"), **out)
        print("c3", out["Q_nnz"], like, out["logdetQ"], out["logdetQc"], "%.0f s" % (time.time() - t0))
    finally:
        rh.set_sparse_factor(None)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    which = set(sys.argv[1:]) or {"c1", "c2", "c3"}
    if "c1" in which:
        make_c1()
    if "c2" in which:
        make_c2()
    if "c3" in which:
        make_c3()
