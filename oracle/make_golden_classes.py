"""Record the host-side surface of every model class of the UNMODIFIED reference (type string, default parameter
vector, getPars(onlySelf=False), the text of print(par)) -> ``tests/golden/classes/surface.json``.
TEST INFRASTRUCTURE ONLY; runs in the build container.

    python oracle/make_golden_classes.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness as rh  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden", "classes")
NAMES = ["whittle-matern", "var-whittle-matern", "advection-diffusion", "advection-var-diffusion", "cov-advection-diffusion",
         "cov-advection-var-diffusion", "var-advection-diffusion", "var-advection-var-diffusion", "seperable-spatial-temporal"]

if __name__ == "__main__":
    sp = rh.load_reference()
    x, y, t = np.linspace(0, 3, 8), np.linspace(0, 2, 7), np.linspace(0, 1, 10)
    rows = []
    for num, name in zip([1, -1, 2, 3, 4, 5, 6, 7, 8], NAMES):
        for ha, ani in ((True, True), (False, True), (False, False)):
            for bc in (1, 3):
                timed = "whittle" not in name
                g = sp.grid(x=x, y=y, t=t) if timed else sp.grid(x=x, y=y)
                mod = sp.model(grid=g, spde=name, ha=ha, anisotropic=ani, bc=bc).mod
                par = np.array(mod.getPars(), dtype="float64")
                row = {"spde": name, "num": num, "ha": ha, "ani": ani, "bc": bc, "type": mod.type, "par": par.tolist()}
                try:
                    row["par_all"] = np.array(mod.getPars(onlySelf=False), dtype="float64").tolist()
                except TypeError:
                    row["par_all"] = None
                try:
                    row["print"] = mod.print(par)
                except Exception as e:      # a few print() bodies index past the vector
                    row["print"] = "ERR:" + type(e).__name__
                # initFit's return value is the start vector of Model.fit (model.py:43-47)
                n = g.n if timed else g.Ns
                idx = np.arange(0, n, 3)
                for fq in ((False, True) if timed and "seperable" not in name else (None,)):
                    m2 = sp.model(grid=g, spde=name, ha=ha, anisotropic=ani, bc=bc).mod
                    kw = {"idx": idx}
                    if fq is not None:
                        kw["fitQ0"] = fq
                    if name.startswith("cov"):
                        kw["ww"] = np.zeros((g.Ns, 4))
                    row["x0_%s" % fq] = np.array(m2.initFit(np.ones((idx.size, 2)), **kw), dtype="float64").tolist()
                rows.append(row)
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, "surface.json"), "w") as f:
        json.dump(rows, f, indent=0)
    print("wrote", len(rows), "rows")
