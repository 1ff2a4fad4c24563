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
