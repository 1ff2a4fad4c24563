"""Host grid helpers on the caller side of the path against the unmodified reference (fixtures written by
oracle/make_golden_grid.py): ``Grid.assimilate_adv`` -- cell-centred currents to the face velocities of the
covariate-driven advection classes, with the extension ramp -- bit for bit."""
import os

import numpy as np

from spdepy_b200.grids import grid

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "grid", "assimilate_adv.npz"))


def test_assimilate_adv_matches_reference():
    for c, (M, N, T, ext) in enumerate(G["cases"]):
        g = grid(x=np.linspace(0, 3, M), y=np.linspace(0, 2, N), t=np.linspace(0, 1, T), extend=None if ext < 0 else int(ext))
        ww = g.assimilate_adv(G["we%d" % c], G["wn%d" % c])
        assert ww.shape == G["ww%d" % c].shape == (g.Ns, 4)
        assert np.array_equal(ww, G["ww%d" % c])


def test_assimilate_adv_constant_field():
    """A constant current stays constant on every face of the interior and is ramped to zero across the extension."""
    g = grid(x=np.linspace(0, 3, 6), y=np.linspace(0, 2, 5), t=np.linspace(0, 1, 2), extend=2)
    ww = g.assimilate_adv(np.full(30, 2.0), np.full(30, -1.0))
    inner = g.obs_nodes()[:30] if hasattr(g, "obs_nodes") else None
    assert np.all((ww[:, 0] >= 0) & (ww[:, 0] <= 2.0)) and np.all((ww[:, 1] <= 0) & (ww[:, 1] >= -1.0))
    if inner is not None:
        assert np.array_equal(ww[inner], np.tile([2.0, -1.0, 2.0, -1.0], (30, 1)))
    assert np.array_equal(ww[0], np.zeros(4))          # outer corner of the extension


def test_transdiff_matches_reference():
    """``transDiff`` of the nine half-angle classes (host helper, no device work) against the unmodified reference,
    per-class quirks included (which of par[0] / par[1] scales tvx, tvy; constant form in VarWhittleMaternHa2D)."""
    import pytest
    import spdepy_b200 as sp
    T = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "grid", "transdiff.npz"))
    names = sorted({k.split("|")[0] for k in T.files})
    assert len(names) == 9
    x, y, t = np.linspace(0, 3, 8), np.linspace(0, 2, 7), np.linspace(0, 1, 10)
    for name in names:
        timed = "whittle" not in name
        g = sp.grid(x=x, y=y, t=t) if timed else sp.grid(x=x, y=y)
        mod = sp.model(grid=g, spde=name, ha=True, bc=3).mod
        assert mod.getPars().shape == T[name + "|par"].shape
        mod.transDiff(T[name + "|par"])
        for k in ("tgamma", "tvx", "tvy"):
            got, ref = np.asarray(getattr(mod, k), dtype="float64"), T[name + "|" + k]
            assert got.shape == ref.shape and np.array_equal(got, ref), (name, k)
    ani = sp.model(grid=sp.grid(x=x, y=y), spde="whittle-matern", ha=False, anisotropic=True, bc=3).mod
    with pytest.raises(AttributeError):
        ani.transDiff()


def test_grid_helpers_match_reference():
    """Both mesh classes: selection matrices (plain, observation subset, with intercept, with a covariate -- the spatial
    mesh of the reference ignores ``scale``), spline-field evaluations, index maps, volume matrices."""
    from scipy import sparse
    H = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "grid", "helpers.npz"))

    def ref(key):
        if key + "|data" in H.files:
            return sparse.csc_matrix((H[key + "|data"], H[key + "|indices"], H[key + "|indptr"]), shape=tuple(H[key + "|shape"]))
        return H[key]

    def same(a, b):
        if sparse.issparse(b):
            a = sparse.csc_matrix(a)
            return a.shape == b.shape and abs(a - b).max() == 0
        a = np.asarray(a)
        return a.shape == b.shape and np.array_equal(a, b)

    for c, (M, N, T, ext) in enumerate(H["cases"]):
        x, y = np.linspace(0, 3, M), np.linspace(0, 2, N)
        t = None if T < 0 else np.linspace(0, 1, T)
        e = None if ext < 0 else int(ext)
        g = grid(x=x, y=y, t=t, extend=e)
        k = "c%d|" % c
        idx = H[k + "idx"]
        p9 = np.random.default_rng(4).normal(size=9)
        got = {"shape": np.array(g.shape), "S": g.getS(), "S_idx": g.getS(idx), "evalB": g.evalB(p9), "evalBH": g.evalBH(p9),
               "Dv": g.Dv, "iDv": g.iDv, "h": np.array([g.hx, g.hy])}
        if T > 0:
            got["evalAdv"] = g.evalAdv(np.random.default_rng(5).normal(size=18))
            got["getIdx"] = np.array([g.getIdx(np.array([1, 2, 1])), g.getIdx(np.array([1, 2, 1]), extend=False)])
            got["dt"] = g.dt
        else:
            got["getIdx"] = np.array([g.getIdx(np.array([1, 2])), g.getIdx(np.array([1, 2]), extend=False)])
        g.addCov(H[k + "cov"], scale=True)
        got["S_cov"], got["S_cov_idx"] = g.getS(), g.getS(idx)
        g2 = grid(x=x, y=y, t=t, extend=e)
        g2.addInt()
        got["S_int_idx"] = g2.getS(idx)
        for name, v in got.items():
            assert same(v, ref(k + name)), (c, name)
        # obs_nodes (this package's device-side form of S) addresses exactly the unit entries of S
        S = sparse.csc_matrix(ref(k + "S_idx"))
        assert np.array_equal(S.tocsr().indices, grid(x=x, y=y, t=t, extend=e).obs_nodes(idx))
