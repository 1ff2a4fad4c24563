"""Host-side surface of all 27 model classes x {bc 1, 3} against the unmodified reference (fixture written by
oracle/make_golden_classes.py): class selection by name and by number (spdes/__init__.py), the type string, the default
parameter vector, getPars(onlySelf=False) with the initial-field block, the set/get round trip and the text of
print(par).  Constructing a model does no device work, so this runs without a GPU."""
import json
import os

import numpy as np
import pytest

import spdepy_b200 as sp

ROWS = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "classes", "surface.json")))
X, Y, T = np.linspace(0, 3, 8), np.linspace(0, 2, 7), np.linspace(0, 1, 10)


@pytest.mark.parametrize("row", ROWS, ids=lambda r: "%s-ha%d-ani%d-bc%d" % (r["spde"], r["ha"], r["ani"], r["bc"]))
def test_surface(row):
    timed = "whittle" not in row["spde"]
    g = sp.grid(x=X, y=Y, t=T) if timed else sp.grid(x=X, y=Y)
    mod = sp.model(grid=g, spde=row["spde"], ha=row["ha"], anisotropic=row["ani"], bc=row["bc"]).mod
    assert mod.type == row["type"]
    par = np.array(mod.getPars(), dtype="float64")
    assert np.array_equal(par, np.array(row["par"]))
    if row["par_all"] is not None:
        assert np.array_equal(np.array(mod.getPars(onlySelf=False), dtype="float64"), np.array(row["par_all"]))
    if not row["print"].startswith("ERR:"):
        assert mod.print(par) == row["print"]
    # numeric model id selects the same class
    g2 = sp.grid(x=X, y=Y, t=T) if timed else sp.grid(x=X, y=Y)
    assert sp.model(grid=g2, spde=row["num"], ha=row["ha"], anisotropic=row["ani"], bc=row["bc"]).mod.type == row["type"]
    # set / get round trip on a perturbed vector
    p2 = par + 0.1 * np.random.default_rng(0).normal(size=par.size)
    mod.setPars(p2)
    assert np.array_equal(np.array(mod.getPars(), dtype="float64"), p2)


@pytest.mark.parametrize("row", ROWS[1::2], ids=lambda r: "%s-ha%d-ani%d-bc%d" % (r["spde"], r["ha"], r["ani"], r["bc"]))
def test_initfit_start_vector(row, monkeypatch):
    """The vector initFit returns is where Model.fit starts (model.py:43-47): per class, with and without the joint
    initial-field block (fitQ0).  The device copies of the observation tables are stubbed out, nothing else is."""
    import spdepy_b200.spdes.base as base
    import spdepy_b200.spdes.separable as sep
    monkeypatch.setattr(base, "to_dev", lambda a, dt=None: a)
    monkeypatch.setattr(sep, "to_dev", lambda a, dt=None: a)
    timed = "whittle" not in row["spde"]
    g = sp.grid(x=X, y=Y, t=T) if timed else sp.grid(x=X, y=Y)
    idx = np.arange(0, g.n if timed else g.Ns, 3)
    for key in [k for k in row if k.startswith("x0_")]:
        fq = {"None": None, "True": True, "False": False}[key[3:]]
        mod = sp.model(grid=g, spde=row["spde"], ha=row["ha"], anisotropic=row["ani"], bc=row["bc"]).mod
        kw = {"idx": idx}
        if fq is not None:
            kw["fitQ0"] = fq
        if row["spde"].startswith("cov"):
            kw["ww"] = np.zeros((g.Ns, 4))
        x0 = np.array(mod.initFit(np.ones((idx.size, 2)), **kw), dtype="float64")
        assert np.array_equal(x0, np.array(row[key])), key
        assert mod.r == 2 and mod.S.shape[0] == idx.size
