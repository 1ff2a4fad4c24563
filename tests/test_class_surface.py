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
