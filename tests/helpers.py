"""Shared helpers for the parity tests: load a golden fixture and rebuild its inputs."""
import glob
import os

import numpy as np
from scipy import sparse

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz")))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    d = {k: z[k] for k in z.files}
    for k in ("spde", "mod0_spde", "type"):
        d[k] = str(d[k])
    for k in ("ha", "ani", "fitQ0"):
        d[k] = bool(d[k])
    for k in ("bc", "M", "N", "T", "ext", "nh1"):
        d[k] = int(d[k])
    d["T"] = None if d["T"] < 0 else d["T"]
    d["ext"] = None if d["ext"] < 0 else d["ext"]
    d["like"] = float(d["like"])
    n = d["Q_indptr"].size - 1
    d["Q"] = sparse.csc_matrix((d["Q_data"], d["Q_indices"], d["Q_indptr"]), shape=(n, n))
    return d


def spec_key(spde, ha, ani):
    """reference class selection (spdes/__init__.py:1-105) -> oracle spec key"""
    mid = {"whittle-matern": "whittle-matern%s-2D", "var-whittle-matern": "var-whittle-matern%s-2D"}
    if spde in mid:
        return mid[spde] % ("-ha" if ha else ("-anisotropic" if ani else "-isotropic"))
    head, tail = spde.rsplit("diffusion", 1)
    return head + ("ha-diffusion" if ha else ("diffusion" if ani else "idiffusion")) + "-2D"


def canon(Q):
    Q = sparse.csc_matrix(Q).copy()
    Q.sum_duplicates()
    Q.eliminate_zeros()
    Q.sort_indices()
    return Q


def make_grids(d):
    from spdepy_b200.grids import grid
    g = grid(x=d["x"], y=d["y"], t=None if d["T"] is None else d["t"], extend=d["ext"])
    g0 = grid(x=d["x"], y=d["y"], extend=d["ext"]) if d["T"] is not None else None
    return g, g0


def make_oracle(d):
    import spde_oracle as so
    g, g0 = make_grids(d)
    if d["spde"] == "seperable-spatial-temporal":
        return so.OracleSeparable(g, bc=d["bc"], variant="ha" if d["ha"] else ("ani" if d["ani"] else "iso"))
    mod0 = None
    if g0 is not None:
        mod0 = so.OracleSPDE(spec_key(d["mod0_spde"], d["ha"], d["ani"]), g0, bc=d["bc"], par=d["mod0_par"])
    ww = d["ww"] if "ww" in d and d["ww"].size else None
    mod = so.OracleSPDE(spec_key(d["spde"], d["ha"], d["ani"]), g, mod0=mod0, bc=d["bc"], ww=ww)
    return mod


def relerr(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)
