"""The oracle's CPU supernodal Cholesky (stand-in for the absent CHOLMOD) against dense LAPACK."""
import time

import numpy as np
import pytest
from scipy import sparse

import cpu_cholesky as cc
import spde_oracle as so
from helpers import load_golden, make_oracle
from spdepy_b200 import _lib
from spdepy_b200.pattern import Pattern


@pytest.mark.parametrize("name", ["wm_ani_bc1_ext", "ad_ani_bc3_q0", "vavd_ani_bc2"])
def test_supernodal_factor_vs_dense(name):
    d = load_golden(name)
    mod = make_oracle(d)
    mod.setQ(d["par"])
    M, N = mod.grid.shape[0], mod.grid.shape[1]
    plan = _lib.PlanHandle(M, N, mod.grid.T if mod.spec.timed else 1, d["bc"])
    f = cc.SupernodalFactor(mod.Q, plan=plan)
    g = so.DenseFactor(mod.Q, perm=plan.perm.astype(np.int64))
    assert abs(f.logdet() - g.logdet()) < 1e-11 * abs(g.logdet())
    B = np.random.default_rng(0).normal(size=(plan.n, 3))
    assert np.abs(f.solve_A(B) - g.solve_A(B)).max() < 1e-10 * np.abs(g.solve_A(B)).max()
    assert np.abs(f.solve_Lt(B) - g.solve_Lt(B)).max() < 1e-10 * np.abs(g.solve_Lt(B)).max()
    assert np.abs(f.solve_A(B[:, 0]) - g.solve_A(B[:, 0])).max() < 1e-10


def test_oracle_loglike_with_supernodal_factor_matches_golden():
    d = load_golden("ad_ani_bc3_q0")
    mod = make_oracle(d)
    M, N = mod.grid.shape[0], mod.grid.shape[1]
    plan = _lib.PlanHandle(M, N, mod.grid.T, d["bc"])
    so.set_factor(cc.factor_with_plan(plan))
    try:
        mod.initFit(d["data"], idx=d["idx"])
        like, jac = mod.logLike(d["par"], grad=True, probes=d["probes"].astype(np.int64))
    finally:
        so.set_factor(None, None)
    assert abs(like - d["like"]) < 1e-9 * abs(d["like"])
    assert np.abs(jac - d["jac"]).max() < 1e-9 * np.abs(d["jac"]).max()
