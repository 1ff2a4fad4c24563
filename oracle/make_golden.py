"""Generate ``tests/golden/*.npz`` from the UNMODIFIED reference (run in the build container only).

    python oracle/make_golden.py            # rewrites every fixture

Each fixture holds the inputs (mesh, model, theta, observation indices, data, probes) and what the
reference itself returned for them through ``oracle/ref_harness.py``: ``Q`` (canonical CSC),
``logLike`` value and gradient with the probes fixed by ``np.random.seed``, ``mu_c`` and the two
log-determinants, one ``Model.sample`` block (identity permutation in the stand-in factoriser)
and one ``Model.update``.  The factoriser behind the reference is the dense LAPACK stand-in for
the absent CHOLMOD, so the quantities that do not depend on the
permutation are exactly what the reference computes.
"""
from __future__ import annotations

import os
import sys

import numpy as np
from scipy import sparse

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness as rh  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

# name, spde, ha, ani, bc, (M, N, T), extend, mod0 spde, fitQ0
CASES = [
    ("wm_iso_bc3", "whittle-matern", False, False, 3, (12, 10, None), None, None, False),
    ("wm_ani_bc1_ext", "whittle-matern", False, True, 1, (12, 10, None), 2, None, False),
    ("wm_ha_bc3", "whittle-matern", True, True, 3, (11, 9, None), None, None, False),
    ("varwm_ani_bc1_ext", "var-whittle-matern", False, True, 1, (12, 10, None), 2, None, False),
    ("varwm_ani_bc2", "var-whittle-matern", False, True, 2, (9, 8, None), None, None, False),
    ("ad_ani_bc3_q0", "advection-diffusion", False, True, 3, (10, 8, 4), None, "whittle-matern", True),
    ("ad_ani_bc1", "advection-diffusion", False, True, 1, (10, 8, 4), None, "whittle-matern", False),
    ("ad_iso_bc3_ext", "advection-diffusion", False, False, 3, (8, 7, 3), 2, "whittle-matern", True),
    ("ad_ha_bc1_q0", "advection-diffusion", True, True, 1, (9, 8, 3), None, "whittle-matern", True),
    ("avd_ani_bc3", "advection-var-diffusion", False, True, 3, (8, 7, 3), None, "whittle-matern", False),
    ("vad_ani_bc1_ext", "var-advection-diffusion", False, True, 1, (8, 7, 3), 2, "whittle-matern", True),
    ("vavd_ani_bc1_ext_q0", "var-advection-var-diffusion", False, True, 1, (8, 7, 3), 2, "var-whittle-matern", True),
    ("vavd_ani_bc3", "var-advection-var-diffusion", False, True, 3, (8, 7, 3), None, "var-whittle-matern", False),
    ("vavd_ani_bc2", "var-advection-var-diffusion", False, True, 2, (8, 7, 3), None, "var-whittle-matern", False),
    ("avhd_bc3", "advection-var-diffusion", True, True, 3, (8, 7, 3), None, "whittle-matern", False),
    ("vahd_bc1_ext_q0", "var-advection-diffusion", True, True, 1, (8, 7, 3), 2, "whittle-matern", True),
    ("vavhd_bc3", "var-advection-var-diffusion", True, True, 3, (8, 7, 3), None, "whittle-matern", False),
    ("cad_ani_bc1_q0", "cov-advection-diffusion", False, True, 1, (8, 7, 3), None, "whittle-matern", True),
    ("cavd_ha_bc3", "cov-advection-var-diffusion", True, True, 3, (8, 7, 3), None, "whittle-matern", False),
    ("vavd_iso_bc3_q0", "var-advection-var-diffusion", False, False, 3, (8, 7, 3), None, "whittle-matern", True),
    # var-Whittle-Matern classes without makeQ in the reference (inline assembly in logLike)
    ("varwm_iso_bc3", "var-whittle-matern", False, False, 3, (11, 9, None), None, None, False),
    ("varwm_ha_bc1_ext", "var-whittle-matern", True, True, 1, (10, 9, None), 2, None, False),
    # separable space-time model Q = Qt (x) Qs (the reference ignores mod0 for this family)
    ("sep_ani_bc3", "seperable-spatial-temporal", False, True, 3, (8, 7, 4), None, "whittle-matern", False),
    ("sep_ani_bc1_ext", "seperable-spatial-temporal", False, True, 1, (7, 6, 3), 1, "whittle-matern", False),
    # half-angle / isotropic separable classes: Qt = sigma * tridiag(-a, 1 + a^2, -a) is built with a hard-coded
    # range(10) (seperable_spatial_temporal_ha2D.py:205-224), so they are only defined for T = 10
    ("sep_ha_bc3", "seperable-spatial-temporal", True, True, 3, (7, 6, 10), None, "whittle-matern", False),
    ("sep_iso_bc1_ext", "seperable-spatial-temporal", False, False, 1, (6, 5, 10), 1, "whittle-matern", False),
]


def canon(Q):
    Q = sparse.csc_matrix(Q).copy()
    Q.sum_duplicates()
    Q.eliminate_zeros()
    Q.sort_indices()
    return Q


def theta_for(mod, rng, fitQ0):
    """A generic, well-conditioned parameter vector: the class default perturbed."""
    par = np.array(mod.getPars(onlySelf=not fitQ0) if "onlySelf" in mod.getPars.__code__.co_varnames else mod.getPars(),
                   dtype="float64")
    par = par + 0.3 * rng.normal(size=par.size)
    par[-1] = np.log(50.0)
    return par


def run_case(case):
    name, spde, ha, ani, bc, (M, N, T), ext, mod0_spde, fitQ0 = case
    sp = rh.load_reference()
    rng = np.random.default_rng(sum(ord(c) for c in name))
    x = np.linspace(0.0, 15.0 * (M - 1) / 49.0, M)
    y = np.linspace(0.0, 15.0 * (N - 1) / 49.0, N)
    t = None if T is None else np.linspace(0.0, 2.0 * (T - 1) / 19.0, T)
    g = sp.grid(x=x, y=y, t=t, extend=ext)
    kw = {}
    mod0_par = None
    if T is not None:
        g0 = sp.grid(x=x, y=y, extend=ext)
        m0 = sp.model(grid=g0, spde=mod0_spde, ha=ha, anisotropic=ani, bc=bc)
        mod0_par = np.array(m0.mod.getPars(), dtype="float64")
        mod0_par = mod0_par + 0.2 * rng.normal(size=mod0_par.size)
        m0.mod.setQ(par=mod0_par)
        kw["mod0"] = m0
    mod = sp.model(grid=g, spde=spde, ha=ha, anisotropic=ani, bc=bc, **kw)
    ww = np.zeros((0, 4))
    if spde.startswith("cov-"):
        # covariate-driven advection: the face velocities come from outside (cov_advection_diffusion2D.py:50-60)
        ww = rng.normal(size=(g.Ns, 4))
        mod.mod.ww = ww
        mod.mod.dA_w = mod.mod.Aw(ww)
    par = theta_for(mod.mod, rng, fitQ0)
    rh.set_permutation(None)
    mod.mod.setQ(par=par)
    Q = canon(mod.mod.Q)
    mod.setModel()
    X = mod.sample(n=4, seed=3, simple=True)
    nprod = M * N * (T or 1)
    nobs = nprod // 3
    idx = np.sort(rng.choice(nprod, nobs, replace=False))
    data = X[idx, :2] + 0.1 * rng.normal(size=(nobs, 2))
    if spde.startswith("cov-"):
        mod.mod.initFit(data, idx=idx, fitQ0=fitQ0, ww=ww)
    elif T is not None:
        mod.mod.initFit(data, idx=idx, fitQ0=fitQ0)
    else:
        mod.mod.initFit(data, idx=idx)
    nh1 = 16
    probes = rh.seeded_probes(g.n, nh1, 4)
    like, jac = rh.loglike_seeded(mod.mod, par, nh1=nh1, grad=True, seed=4)
    like0 = rh.loglike_seeded(mod.mod, par, nh1=nh1, grad=False, seed=4)
    assert like0 == like
    # intermediate quantities (recomputed exactly as logLike does, advection_diffusion2D.py:190-198)
    from sksparse.cholmod import cholesky
    tau = np.exp(par[-1])
    if hasattr(mod.mod, "makeQ"):
        res = mod.mod.makeQ(par=par, grad=False)          # (Q, Q_fac, None) or, for the separable class, (Q, Q_fac)
        Qm, Qf = res[0], res[1]
    else:                                                 # var-whittle-matern iso / ha: setQ is the only assembly entry
        S_obs = mod.mod.S                                 # setQ resets S to the full selection matrix
        mod.mod.setQ(par=par)
        Qm, Qf = mod.mod.Q, mod.mod.Q_fac
        mod.mod.S = S_obs
    S = mod.mod.S
    Qc = Qm + S.T @ S * tau
    Qcf = cholesky(Qc)
    mu_c = Qcf.solve_A(S.T @ data * tau)
    # one conditioning step (model.py:120-127)
    mod.mod.setQ(par=par)
    mod.setModel()
    mod.update(y=data[:, 0], idx=idx)
    out = dict(
        spde=spde, ha=ha, ani=ani, bc=bc, M=M, N=N, T=-1 if T is None else T, ext=-1 if ext is None else ext,
        mod0_spde="" if mod0_spde is None else mod0_spde, fitQ0=fitQ0,
        x=x, y=y, t=np.zeros(0) if t is None else t, par=par,
        mod0_par=np.zeros(0) if mod0_par is None else mod0_par,
        type=mod.mod.type, ww=ww,
        Q_data=Q.data, Q_indices=Q.indices.astype(np.int32), Q_indptr=Q.indptr.astype(np.int32),
        sample=X, idx=idx, data=data, nh1=nh1, probes=probes.astype(np.int8),
        like=like, jac=jac, mu_c=mu_c, logdetQ=Qf.logdet(), logdetQc=Qcf.logdet(),
        upd_mu=mod.mu, upd_Qdiag=mod.Q.diagonal(),
    )
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    return name, Q.shape[0], Q.nnz, like


def copy_fitted_theta():
    """The 92 fitted parameters of the SINMOD application (SURVEY.md section 8d, C3 inputs): a data fixture of
    the reference (``examples/sinmod_example/fits/var_advection_var_diffusion_ani_bc1.npy``)."""
    p = np.load(os.path.join(rh.REF_ROOT, "examples", "sinmod_example", "fits", "var_advection_var_diffusion_ani_bc1.npy"))
    np.save(os.path.join(OUT, "c3_theta.npy"), p)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    copy_fitted_theta()
    only = set(sys.argv[1:])
    for c in CASES:
        if only and c[0] not in only:
            continue
        print(*run_case(c))
