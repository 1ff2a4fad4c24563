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
