"""GPU parity of the numeric hot path: factorisation / logdet / solves / samples / selected inverse /
log-likelihood and gradient, against the reference's golden outputs and the CPU oracle.
Tolerances (BASELINE.json north_star): 1e-9 relative in FP64."""
import numpy as np
import pytest
import torch
from scipy import sparse

import spde_oracle as so
from helpers import golden_names, load_golden, make_grids, make_oracle, relerr

pytestmark = pytest.mark.gpu


def _build(d):
    import spdepy_b200 as sp
    g, g0 = make_grids(d)
    kw = {}
    if g0 is not None:
        kw["mod0"] = sp.model(grid=g0, spde=d["mod0_spde"], ha=d["ha"], anisotropic=d["ani"], bc=d["bc"], parameters=d["mod0_par"])
    mod = sp.model(grid=g, spde=d["spde"], ha=d["ha"], anisotropic=d["ani"], bc=d["bc"], **kw)
    if "ww" in d and d["ww"].size:
        mod.mod.ww = d["ww"]
    return mod


def _dense(Q):
    A = sparse.tril(Q).toarray()
    return A + np.tril(A, -1).T


@pytest.mark.parametrize("name", golden_names())
def test_factor_methods(name):
    d = load_golden(name)
    mod = _build(d)
    mod.mod.setQ(d["par"])
    fac = mod.mod.Q_fac
    A = _dense(d["Q"])
    n = A.shape[0]
    perm = fac.P().astype(np.int64)
    L = np.linalg.cholesky(A[np.ix_(perm, perm)])
    assert abs(fac.logdet() - 2 * np.log(np.diag(L)).sum()) <= 1e-9 * abs(fac.logdet())
    assert abs(fac.logdet() - float(d["logdetQ"])) <= 1e-9 * abs(float(d["logdetQ"]))
    rng = np.random.default_rng(0)
    B = rng.normal(size=(n, 5))
    assert relerr(fac.solve_A(B), np.linalg.solve(A, B)) < 1e-9
    assert relerr(fac.solve_A(B[:, 0]), np.linalg.solve(A, B[:, 0])) < 1e-9
    x = fac.apply_Pt(fac.solve_Lt(B, use_LDLt_decomposition=False))
    ref = np.empty_like(B)
    ref[perm] = np.linalg.solve(L.T, B)
    assert relerr(x, ref) < 1e-9
    assert relerr(fac.solve_L(fac.apply_P(B)), np.linalg.solve(L, B[perm])) < 1e-9


@pytest.mark.parametrize("name", golden_names())
def test_sample_update_against_oracle(name):
    """Model.sample / Model.update with identical normal draws and the build's permutation."""
    d = load_golden(name)
    mod = _build(d)
    mod.mod.setQ(d["par"])
    mod.setModel()
    X = mod.sample(n=4, seed=3, simple=True)
    perm = mod.Q_fac.P().astype(np.int64)
    so.set_factor(None, lambda n: perm)
    try:
        Xo = so.sample(d["Q"], mod.grid.getS(), n=4, seed=3)
    finally:
        so.set_factor(None, None)
    assert relerr(X, Xo) < 1e-9
    # a column block of the same draw (the shard of one rank, parallel.sample_sharded)
    assert relerr(mod.sample(n=4, seed=3, simple=True, cols=slice(1, 3)), Xo[:, 1:3]) < 1e-9
    mod.update(y=d["data"][:, 0], idx=d["idx"])
    assert relerr(mod.mu, d["upd_mu"]) < 1e-9
    assert relerr(mod.getQ().diagonal(), d["upd_Qdiag"]) < 1e-13


@pytest.mark.parametrize("name", golden_names())
def test_selected_inverse(name):
    d = load_golden(name)
    mod = _build(d)
    mod.mod.setQ(d["par"])
    eng = mod.mod.engine
    Z = eng.selinv(0).cpu().numpy()
    Zd = np.linalg.inv(_dense(d["Q"]))
    full = eng.pattern.to_csc(Z).toarray()
    mask = eng.pattern.to_csc(np.ones_like(Z)).toarray() != 0
    assert np.abs(full[mask] - Zd[mask]).max() <= 1e-9 * np.abs(Zd).max()
    mod.setModel()
    assert relerr(mod.qinv(simple=False), np.diag(Zd)) < 1e-9


@pytest.mark.parametrize("name", golden_names())
def test_loglike_and_hutchinson_gradient_vs_reference(name):
    """like and jac with the reference's own probe draw (np.random.seed(4)) against the values the
    unmodified reference returned."""
    d = load_golden(name)
    mod = _build(d)
    m = mod.mod
    m.initFit(d["data"], idx=d["idx"], fitQ0=d["fitQ0"])
    like, jac = m.logLike(d["par"], grad=True, probes=d["probes"].astype(np.float64))
    assert abs(like - d["like"]) <= 1e-9 * abs(d["like"])
    assert relerr(m.last["mu_c"].cpu().numpy(), d["mu_c"].reshape(-1, 2)) < 1e-9
    assert abs(m.last["logdetQ"] - float(d["logdetQ"])) <= 1e-9 * abs(float(d["logdetQ"]))
    assert abs(m.last["logdetQc"] - float(d["logdetQc"])) <= 1e-9 * abs(float(d["logdetQc"]))
    assert np.abs(jac - d["jac"]).max() <= 1e-9 * np.abs(d["jac"]).max(), (jac, d["jac"])
    assert m.logLike(d["par"], grad=False) == pytest.approx(like, rel=1e-12)


@pytest.mark.parametrize("name", golden_names())
def test_exact_gradient_vs_oracle_dense_inverse(name):
    """Takahashi gradient against the oracle's formula evaluated with dense inverses."""
    d = load_golden(name)
    mod = _build(d)
    m = mod.mod
    m.initFit(d["data"], idx=d["idx"], fitQ0=d["fitQ0"])
    like, jac = m.logLike(d["par"], grad=True, exact_grad=True)
    orc = make_oracle(d)
    orc.initFit(d["data"], idx=d["idx"])
    like_o, jac_o = orc.logLike_exact(d["par"])
    assert abs(like - like_o) <= 1e-9 * abs(like_o)
    assert np.abs(jac - jac_o).max() <= 1e-9 * np.abs(jac_o).max(), (jac, jac_o)


@pytest.mark.parametrize("name", ["ad_ani_bc3_q0", "vavd_ani_bc3", "wm_iso_bc3"])
def test_device_resident_scalars_read_once(name):
    """The scalars of one logLike are collected on the device and read back in ONE copy (engine.ScalarPool): the
    device-to-host traffic of an evaluation is the pool (a few hundred doubles) plus the three factorisation status words,
    and the result equals the eager path (every reduction copied and synchronised on its own) to the last bit or two."""
    from spdepy_b200.engine import COUNTERS
    d = load_golden(name)
    mod = _build(d)
    m = mod.mod
    m.initFit(d["data"], idx=d["idx"], fitQ0=d["fitQ0"])
    m.logLike(d["par"], grad=True, exact_grad=True)               # warm-up (schedules, graphs)
    COUNTERS["d2h"] = 0
    like, jac = m.logLike(d["par"], grad=True, exact_grad=True)
    lazy_bytes = COUNTERS["d2h"]
    assert lazy_bytes <= 8 * (d["par"].size + 64)
    COUNTERS["d2h"] = 0
    m._pool = None
    like_e, jac_e = m._logLike(d["par"], 100, True, None, True)   # eager: host floats from every reduction
    assert COUNTERS["d2h"] >= 8 * 8
    assert like == pytest.approx(like_e, rel=1e-13)
    assert np.abs(jac - jac_e).max() <= 1e-12 * np.abs(jac_e).max()


def test_not_positive_definite_raises():
    import spdepy_b200 as sp
    from spdepy_b200._lib import NotPositiveDefiniteError
    d = load_golden("wm_iso_bc3")
    mod = _build(d)
    mod.mod.setQ(d["par"])
    eng = mod.mod.engine
    Q = mod.mod._state["Q"].clone()
    Q[(eng.nslots // 2) * eng.n + 3] = -1.0
    with pytest.raises(NotPositiveDefiniteError):
        eng.factorize(0, Q)


@pytest.mark.parametrize("name", ["wm_ani_bc1_ext", "varwm_ani_bc2", "ad_ani_bc3_q0", "ad_ha_bc1_q0", "vavd_ani_bc1_ext_q0",
                                  "sep_ani_bc3", "sep_ha_bc3", "sep_iso_bc1_ext"])
def test_lazy_dQ_operators_vs_oracle(name):
    """makeQ(grad=True) returns dQ as lazy operators; dQ[i] @ X must equal the reference's explicit matrices."""
    d = load_golden(name)
    mod = _build(d)
    Q, fac, dQ = mod.mod.makeQ(d["par"], grad=True)
    orc = make_oracle(d)
    Qo, _, dQo = orc.makeQ(d["par"], grad=True)
    assert len(dQ) == len(dQo) == d["par"].size - 1
    X = np.random.default_rng(0).normal(size=(Q.shape[0], 3))
    for i in range(len(dQ)):
        ref = dQo[i] @ X
        got = dQ[i] @ X
        assert np.abs(got - ref).max() <= 1e-9 * max(np.abs(ref).max(), 1e-300), (i, np.abs(got - ref).max(), np.abs(ref).max())
    assert abs(dQ[0].tocsc() - dQo[0]).max() <= 1e-9 * abs(dQo[0]).max()


@pytest.mark.parametrize("name,two", [("wm_ani_bc1_ext", True), ("ad_ani_bc3_q0", False)])
def test_bordered_model_with_regression_columns(name, two):
    """Model.setModel(useCov=True) / update / sample (model.py:73-87,120-153): the latent vector gains regression
    coefficients and the precision a dense border.  Checked against a dense solve of the bordered system and a dense
    Cholesky under the build's permutation (border ordered last)."""
    d = load_golden(name)
    mod = _build(d)
    mod.mod.setQ(d["par"])
    g = mod.grid
    nobs_all = g.M * g.N * g.T
    rng = np.random.default_rng(11)
    cov = rng.uniform(0.5, 2.0, size=nobs_all)
    if two:
        mod.setModel(mu=cov, sigmas=np.log(np.array([0.01, 140.0])), useCov=True, scale=True)
        k = 2
    else:
        mod.setModel(mu=np.zeros(nobs_all), sigmas=np.log(0.5), useCov=True)
        k = 1
    n = mod.mod.engine.n
    S = g.getS()
    assert S.shape == (nobs_all, n + k) and mod.mu.shape == (n + k,)
    Q0 = mod.getQ().toarray()
    assert Q0.shape == (n + k, n + k)
    mu0 = mod.mu.copy()
    idx = d["idx"][: max(10, d["idx"].size // 2)]
    y = rng.normal(size=idx.size)
    tau = 3.0
    mod.update(y=y, idx=idx, tau=tau)
    Si = g.getS(idx).toarray()
    Qd = Q0 + tau * Si.T @ Si
    mu_ref = mu0 + np.linalg.solve(Qd, Si.T @ (y - Si @ mu0)) * tau
    assert relerr(mod.mu, mu_ref) < 1e-9
    assert relerr(mod.getQ().toarray(), Qd) < 1e-12
    # samples: x = P_aug^T L_aug^-T z + mu with the border ordered last
    X = mod.sample(n=3, seed=5, simple=True)
    perm = np.concatenate([mod.mod.engine.plan.perm.astype(np.int64), n + np.arange(k)])
    L = np.linalg.cholesky(Qd[np.ix_(perm, perm)])
    z = np.random.default_rng(5).normal(size=(n + k) * 3).reshape(n + k, 3)
    x = np.empty_like(z)
    x[perm] = np.linalg.solve(L.T, z)
    ref = S.toarray() @ (x + mod.mu[:, None])
    assert relerr(X, ref) < 1e-9
    # marginal variances of [x; beta]: selected inverse of the sparse block + Schur complement of the border
    mvar = mod.qinv(simple=False)
    assert mvar.shape == (n + k,) and relerr(mvar, np.diag(np.linalg.inv(Qd))) < 1e-9


EDGE_CASES = [
    # spde, mod0 spde, ha, ani, bc, M, N, T, ext, r, repeated observations
    ("advection-diffusion", "whittle-matern", False, True, 3, 5, 4, 2, None, 3, True),      # two time slices, ragged mesh
    ("advection-diffusion", "whittle-matern", False, False, 1, 6, 7, 3, 2, 1, False),       # isotropic, Neumann, extension
    ("var-advection-var-diffusion", "var-whittle-matern", False, True, 2, 7, 6, 3, None, 2, True),   # periodic
    ("whittle-matern", None, True, True, 3, 9, 5, None, None, 4, True),                      # spatial, half-angle
    ("advection-var-diffusion", "whittle-matern", True, True, 3, 6, 6, 4, 1, 1, False),
]


@pytest.mark.parametrize("case", EDGE_CASES, ids=lambda c: "%s-bc%d-%dx%dx%s" % (c[0], c[4], c[5], c[6], c[7]))
def test_edge_meshes_against_oracle(case):
    """Small, ragged and minimal meshes (T = 2, M != N, periodic wrap on the smallest legal size), replicated data
    (r > 1), repeated observation indices (S^T S holds counts) and a single observation: likelihood, exact gradient,
    Hutchinson gradient with injected probes and the conditional mean against the CPU oracle."""
    import spdepy_b200 as sp
    spde, spde0, ha, ani, bc, M, N, T, ext, r, rep = case
    x, y = np.linspace(0.0, 3.0, M), np.linspace(0.0, 2.5, N)
    t = None if T is None else np.linspace(0.0, 0.6, T)
    d = {"x": x, "y": y, "t": t, "T": T, "ext": ext, "spde": spde, "mod0_spde": spde0, "ha": ha, "ani": ani, "bc": bc}
    g, g0 = make_grids(d)
    kw = {}
    if g0 is not None:
        m0 = sp.model(grid=g0, spde=spde0, ha=ha, anisotropic=ani, bc=bc)
        d["mod0_par"] = m0.mod.getPars()
        kw["mod0"] = m0
    mod = sp.model(grid=g, spde=spde, ha=ha, anisotropic=ani, bc=bc, **kw)
    m = mod.mod
    rng = np.random.default_rng(3)
    nall = M * N * (T or 1)
    for idx in (np.sort(rng.choice(nall, nall // 2, replace=rep)), np.array([nall // 3])):
        data = rng.normal(size=(idx.size, r))
        par = m.initFit(data, idx=idx, fitQ0=False)
        par = par + 0.05 * rng.normal(size=par.size)
        orc = make_oracle(d)
        orc.initFit(data, idx=idx)
        like_o, jac_o = orc.logLike_exact(par)
        like, jac = m.logLike(par, grad=True, exact_grad=True)
        assert abs(like - like_o) <= 1e-9 * abs(like_o), (like, like_o)
        assert np.abs(jac - jac_o).max() <= 1e-9 * np.abs(jac_o).max(), (jac, jac_o)
        probes = 2.0 * rng.integers(0, 2, size=(g.n, 8)) - 1.0
        like_h, jac_h = m.logLike(par, grad=True, probes=probes)
        like_ho, jac_ho = orc.logLike(par, nh1=8, grad=True, probes=probes)
        assert abs(like_h - like_ho) <= 1e-9 * abs(like_ho)
        assert np.abs(jac_h - jac_ho).max() <= 1e-9 * np.abs(jac_ho).max(), (jac_h, jac_ho)
