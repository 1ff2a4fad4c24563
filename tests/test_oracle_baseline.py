"""The oracle restatement against fixtures recorded from the UNMODIFIED reference ON THE BASELINE CONFIGS
(``oracle/make_golden_baseline.py``; inputs and seeds of SURVEY.md section 8d): C1 = Whittle-Matern 30x30 (six cases,
r = 1 and r = 20, nh1 = 100 probes of ``np.random.seed(4)``), C2 = advection-diffusion 50x50x20 with 5 000 observations.
Q must be bit-identical; like / jac / mu_c / samples to 1e-9.  CPU only."""
import glob
import hashlib
import os

import numpy as np
import pytest
from scipy import sparse

import cpu_cholesky as cc
import spde_oracle as so
import symbolic_oracle as syo
from grid_oracle import OracleGrid
from helpers import canon, relerr

BASE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "baseline")


def _digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _probes(n, nh1=100, seed=4):
    np.random.seed(seed)
    return (2 * np.random.randint(1, 3, n * nh1) - 3).reshape(n, nh1)


def c1_names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(BASE, "c1_*.npz")))


@pytest.mark.parametrize("name", c1_names())
def test_c1_oracle_equals_reference(name):
    d = np.load(os.path.join(BASE, name + ".npz"))
    ha, ani, bc = bool(d["ha"]), bool(d["ani"]), int(d["bc"])
    g = OracleGrid(d["x"], d["x"])
    key = "whittle-matern-%s-2D" % ("ha" if ha else ("anisotropic" if ani else "isotropic"))
    orc = so.OracleSPDE(key, g, bc=bc)
    assert orc.type == str(d["type"])
    par = d["par"]
    orc.setQ(par)
    Q = canon(orc.Q)
    assert np.array_equal(Q.indptr, d["Q_indptr"]) and np.array_equal(Q.indices, d["Q_indices"])
    assert np.array_equal(Q.data, d["Q_data"])                     # bit for bit
    # Model.sample(n=100, seed=0): simple=False is the reference's default (model.py:73,85-86) -> observation noise added
    X = so.sample(orc.Q, g.getS(), n=100, seed=0, tau=np.exp(par[-1]), simple=False, nprod=900)    # identity permutation
    assert relerr(X[:, :20], d["sample20"]) < 1e-9
    assert relerr(np.sqrt((X ** 2).sum(axis=0)), d["sample_colnorms"]) < 1e-9
    for r in (1, 20):
        data = d["sample20"][d["idx"], :r]
        orc.initFit(data, idx=d["idx"])
        like, jac = orc.logLike(par, nh1=100, grad=True, probes=_probes(900))
        assert abs(like - float(d["like_r%d" % r])) <= 1e-9 * abs(float(d["like_r%d" % r]))
        assert np.abs(jac - d["jac_r%d" % r]).max() <= 1e-9 * np.abs(d["jac_r%d" % r]).max()
        assert relerr(orc.last["mu_c"], d["mu_c_r%d" % r]) < 1e-9
        assert abs(orc.last["logdetQc"] - float(d["logdetQc_r%d" % r])) <= 1e-9 * abs(float(d["logdetQc_r%d" % r]))
    assert abs(orc.last["logdetQ"] - float(d["logdetQ"])) <= 1e-9 * abs(float(d["logdetQ"]))


def _c2_oracle(d):
    x, t, bc = d["x"], d["t"], int(d["bc"])
    g = OracleGrid(x, x, t)
    o0 = so.OracleSPDE("whittle-matern-isotropic-2D", OracleGrid(x, x), bc=bc, par=np.array([-2.0, -0.5, np.log(10.0)]))
    return so.OracleSPDE("advection-diffusion-2D", g, mod0=o0, bc=bc), g


@pytest.mark.parametrize("bc", [3, 1])
def test_c2_oracle_equals_reference(bc):
    """Q: SHA-256 of the canonical CSC arrays (the 24 MB matrix itself is not committed) -> bit-identical pattern and values;
    like / jac / mu_c with the oracle's own supernodal Cholesky and symbolic analysis."""
    path = os.path.join(BASE, "c2_bc%d.npz" % bc)
    if not os.path.exists(path):
        pytest.skip("fixture not generated")
    d = np.load(path)
    orc, g = _c2_oracle(d)
    par = d["par"]
    sym = {}

    def impl(A, perm=None):
        n = A.shape[0]
        if n not in sym:
            sym[n] = syo.OracleSymbolic(A, syo.nd_perm(50, 50, 20 if n == 50000 else 1, bc))
        return cc.SupernodalFactor(A, plan=sym[n])

    so.set_factor(impl)
    try:
        orc.initFit(d["data"], idx=d["idx"])
        like, jac = orc.logLike(par, nh1=100, grad=True, probes=_probes(50000))
    finally:
        so.set_factor(None, None)
    Q = canon(orc.last["Q"])
    assert Q.nnz == int(d["Q_nnz"])
    assert _digest(Q.indptr.astype(np.int64)) == str(d["Q_sha_indptr"])
    assert _digest(Q.indices.astype(np.int64)) == str(d["Q_sha_indices"])
    assert _digest(Q.data.astype(np.float64)) == str(d["Q_sha_data"])          # bit for bit
    assert abs(like - float(d["like"])) <= 1e-9 * abs(float(d["like"]))
    assert np.abs(jac - d["jac"]).max() <= 1e-9 * np.abs(d["jac"]).max(), (jac, d["jac"])
    mu = orc.last["mu_c"]
    assert relerr(mu[::50], d["mu_c_rows"]) < 1e-9
    assert relerr(np.sqrt((mu ** 2).sum(axis=0)), d["mu_c_colnorms"]) < 1e-9
    assert abs(orc.last["logdetQ"] - float(d["logdetQ"])) <= 1e-9 * abs(float(d["logdetQ"]))
    assert abs(orc.last["logdetQc"] - float(d["logdetQc"])) <= 1e-9 * abs(float(d["logdetQc"]))


def test_oracle_grid_equals_product_grid():
    """grid_oracle.OracleGrid (the reference arm's mesh) against the product's mesh class, which tests/test_grids.py pins
    to the unmodified reference: bit for bit."""
    from spdepy_b200.grids import grid
    for x, y, t in ((800.0 * np.arange(13), 800.0 * np.arange(11), 10.0 * np.arange(4)),
                    (np.linspace(0, 15, 9), np.linspace(0, 12, 7), None)):
        g, gp = OracleGrid(x, y, t), grid(x=x, y=y, t=t)
        for a in ("bs", "bsH") + (("bsA",) if t is not None else ()):
            assert np.array_equal(getattr(g, a), getattr(gp, a))
        assert (g.hx, g.hy, g.V, g.shape) == (gp.hx, gp.hy, gp.V, gp.shape)
        idx = np.array([0, 5, 17])
        assert (g.getS(idx) != gp.getS(idx)).nnz == 0 and (g.getS() != gp.getS()).nnz == 0
        if t is not None:
            p = np.random.default_rng(0).normal(size=18)
            assert g.dt == gp.dt and np.array_equal(g.evalAdv(p), gp.evalAdv(p))


@pytest.mark.parametrize("shape,bc", [((12, 10, 1), 3), ((10, 9, 4), 1), ((9, 8, 3), 2), ((20, 17, 6), 3)])
def test_oracle_symbolic_against_product_and_dense(shape, bc):
    """oracle/symbolic_oracle (general-pattern etree / supernodes in C) against the product's geometric analysis (same
    nested dissection => identical nnz(L) and sum cc^2) and, numerically, against dense LAPACK."""
    from spdepy_b200 import _lib
    M, N, T = shape
    st = syo.OracleSymbolic.__new__(syo.OracleSymbolic)
    import bench
    full = bench._oracle_stats(M, N, T, bc)
    pl = _lib.PlanHandle(M, N, T, bc).stats()
    assert full["nnzL"] == pl["nnzL"] and full["flops"] == pl["flops"]
    rng = np.random.default_rng(0)
    n = M * N * T
    k = np.arange(n)
    A = sparse.random(n, n, density=0.0, format="csc")
    # an SPD matrix on the mesh pattern
    Ns = M * N
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
    B = sparse.csc_matrix((0.1 * rng.normal(size=r.size), (r, c)), shape=(n, n))
    A = (B + B.T + sparse.eye(n) * 12.0).tocsc()
    F = cc.SupernodalFactor(A, plan=syo.OracleSymbolic(A, syo.nd_perm(M, N, T, bc)))
    Ad = A.toarray()
    assert abs(F.logdet() - np.linalg.slogdet(Ad)[1]) <= 1e-12 * abs(F.logdet())
    b = rng.normal(size=(n, 3))
    assert relerr(F.solve_A(b), np.linalg.solve(Ad, b)) < 1e-11
    L = np.linalg.cholesky(Ad[np.ix_(F.perm, F.perm)])
    assert relerr(F.solve_Lt(b), np.linalg.solve(L.T, b)) < 1e-11
