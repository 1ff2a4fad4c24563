"""Pin the CPU oracle (oracle/spde_oracle.py) against what the unmodified reference returned
(tests/golden/*.npz, written by oracle/make_golden.py).  Tolerances from BASELINE.json north_star:
pattern exact, Q 1e-13 relative, like / jac / mu_c / samples 1e-9 relative."""
import numpy as np
import pytest

import spde_oracle as so
from helpers import canon, golden_names, load_golden, make_oracle, relerr


@pytest.mark.parametrize("name", golden_names())
def test_oracle_matches_reference(name):
    d = load_golden(name)
    so.set_factor(None, None)
    mod = make_oracle(d)
    assert mod.type == d["type"]
    mod.setQ(d["par"])
    Q = canon(mod.Q)
    assert np.array_equal(Q.indptr, d["Q"].indptr) and np.array_equal(Q.indices, d["Q"].indices)
    assert np.array_equal(Q.data, d["Q"].data), "oracle Q is meant to be bit-identical to the reference's"
    # Model.sample (model.py:73-87), identity permutation on both sides
    X = so.sample(mod.Q, mod.grid.getS(), n=4, seed=3)
    assert relerr(X, d["sample"]) < 1e-9
    # logLike + Hutchinson gradient with the reference's probe draw
    mod.initFit(d["data"], idx=d["idx"])
    like, jac = mod.logLike(d["par"], grad=True, probes=d["probes"].astype(np.int64))
    assert abs(like - d["like"]) < 1e-9 * abs(d["like"])
    assert relerr(jac, d["jac"]) < 1e-9
    assert relerr(mod.last["mu_c"], d["mu_c"].reshape(mod.last["mu_c"].shape)) < 1e-9
    assert abs(mod.last["logdetQ"] - d["logdetQ"]) < 1e-9 * abs(d["logdetQ"])
    assert abs(mod.last["logdetQc"] - d["logdetQc"]) < 1e-9 * abs(d["logdetQc"])
    assert mod.logLike(d["par"], grad=False) == like
    # Model.update (model.py:120-127)
    mod.setQ(d["par"])
    tau = np.exp(d["par"][-1])
    Q2, mu2 = so.update(mod.Q, np.zeros(mod.Q.shape[0]), mod.grid.getS(d["idx"]), d["data"][:, 0], tau)
    assert relerr(mu2, d["upd_mu"]) < 1e-9
    assert relerr(Q2.diagonal(), d["upd_Qdiag"]) < 1e-13
