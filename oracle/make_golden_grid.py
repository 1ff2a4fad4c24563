"""Record ``Grid.assimilate_adv`` of the UNMODIFIED reference (``grids/spat2Dtemp_regular_mesh.py:277-344``) ->
``tests/golden/grid/assimilate_adv.npz``.  TEST INFRASTRUCTURE ONLY; runs in the build container.

    python oracle/make_golden_grid.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness as rh  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden", "grid")
CASES = [(7, 5, 3, -1), (6, 8, 4, 2), (5, 5, 2, 3), (9, 4, 2, 1)]       # M, N, T, extend (-1: none)

if __name__ == "__main__":
    sp = rh.load_reference()
    os.makedirs(OUT, exist_ok=True)
    rng = np.random.default_rng(12)
    out = {"cases": np.array(CASES)}
    for c, (M, N, T, ext) in enumerate(CASES):
        g = sp.grid(x=np.linspace(0, 3, M), y=np.linspace(0, 2, N), t=np.linspace(0, 1, T), extend=None if ext < 0 else ext)
        we, wn = rng.normal(size=M * N), rng.normal(size=M * N)
        out["we%d" % c], out["wn%d" % c], out["ww%d" % c] = we, wn, g.assimilate_adv(we, wn)
    np.savez_compressed(os.path.join(OUT, "assimilate_adv.npz"), **out)
    print("wrote", len(CASES), "cases")


def transdiff_cases():
    """``transDiff`` of the ten half-angle classes of the unmodified reference at perturbed default parameters."""
    sp = rh.load_reference()
    x, y, t = np.linspace(0, 3, 8), np.linspace(0, 2, 7), np.linspace(0, 1, 10)
    rng = np.random.default_rng(21)
    out = {}
    names = ["whittle-matern", "var-whittle-matern", "advection-diffusion", "advection-var-diffusion", "cov-advection-diffusion",
             "cov-advection-var-diffusion", "var-advection-diffusion", "var-advection-var-diffusion", "seperable-spatial-temporal"]
    for name in names:
        timed = "whittle" not in name
        g = sp.grid(x=x, y=y, t=t) if timed else sp.grid(x=x, y=y)
        mod = sp.model(grid=g, spde=name, ha=True, bc=3).mod
        par = np.array(mod.getPars(), dtype="float64")
        par = par + 0.2 * rng.normal(size=par.size)
        mod.transDiff(par)
        out[name + "|par"] = par
        for k in ("tgamma", "tvx", "tvy"):
            out[name + "|" + k] = np.asarray(getattr(mod, k), dtype="float64")
    np.savez_compressed(os.path.join(OUT, "transdiff.npz"), **out)
    print("wrote transDiff of", len(names), "classes")


if __name__ == "__main__":
    transdiff_cases()


def helper_cases():
    """Selection matrices (plain, with an intercept, with a scaled covariate), spline-field evaluations, index maps and
    volume matrices of both mesh classes of the unmodified reference."""
    from scipy import sparse
    sp = rh.load_reference()
    out = {}
    cases = [(7, 5, 3, -1), (6, 8, 4, 2), (5, 6, -1, -1), (6, 5, -1, 2)]          # M, N, T (-1: spatial), extend (-1: none)
    out["cases"] = np.array(cases)

    def put(key, v):
        if sparse.issparse(v):
            v = sparse.csc_matrix(v)
            v.sort_indices()
            out[key + "|data"], out[key + "|indices"], out[key + "|indptr"], out[key + "|shape"] = v.data, v.indices, v.indptr, np.array(v.shape)
        else:
            out[key] = np.asarray(v)

    for c, (M, N, T, ext) in enumerate(cases):
        x, y = np.linspace(0, 3, M), np.linspace(0, 2, N)
        t = None if T < 0 else np.linspace(0, 1, T)
        e = None if ext < 0 else ext
        g = sp.grid(x=x, y=y, t=t, extend=e)
        n = M * N * (1 if T < 0 else T)
        idx = np.sort(np.random.default_rng(3).choice(n, n // 3, replace=False))
        k = "c%d|" % c
        put(k + "idx", idx)
        put(k + "shape", np.array(g.shape))
        put(k + "S", g.getS())
        put(k + "S_idx", g.getS(idx))
        p9 = np.random.default_rng(4).normal(size=9)
        put(k + "evalB", g.evalB(p9))
        put(k + "evalBH", g.evalBH(p9))
        put(k + "Dv", g.Dv)
        put(k + "iDv", g.iDv)
        put(k + "h", np.array([g.hx, g.hy]))
        if T > 0:
            put(k + "evalAdv", g.evalAdv(np.random.default_rng(5).normal(size=18)))
            put(k + "getIdx", np.array([g.getIdx(np.array([1, 2, 1])), g.getIdx(np.array([1, 2, 1]), extend=False)]))
            put(k + "dt", g.dt)
        else:
            put(k + "getIdx", np.array([g.getIdx(np.array([1, 2])), g.getIdx(np.array([1, 2]), extend=False)]))
        cov = np.random.default_rng(6).uniform(0.5, 2, size=n)
        put(k + "cov", cov)
        g.addCov(cov, scale=True)
        put(k + "S_cov", g.getS())
        put(k + "S_cov_idx", g.getS(idx))
        g2 = sp.grid(x=x, y=y, t=t, extend=e)
        g2.addInt()
        put(k + "S_int_idx", g2.getS(idx))
    np.savez_compressed(os.path.join(OUT, "helpers.npz"), **out)
    print("wrote grid helpers of", len(cases), "meshes")


if __name__ == "__main__":
    helper_cases()
