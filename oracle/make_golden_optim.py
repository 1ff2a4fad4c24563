"""Record the iterates of the UNMODIFIED reference optimiser (``optim/__init__.py``, ``optim/{adam,rmsprop,ada_grad,
ada_delta,sgd}.py``) on a deterministic test function -> ``tests/golden/optim/trajectories.npz``.
TEST INFRASTRUCTURE ONLY; runs in the build container (needs /root/reference).

    python oracle/make_golden_optim.py
"""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness as rh  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden", "optim")

A = np.array([1.0, 0.3, 2.5, 0.7, 1.3])
C = np.array([0.5, -1.0, 2.0, 0.0, -0.25])


def fun(x):
    """Smooth, non-separable; the fourth gradient component is exactly zero at the start (the rmsprop zero rule)."""
    d = x - C
    f = float((A * d ** 2).sum() + 0.1 * np.sin(x[0] * x[1]))
    g = 2 * A * d
    g[0] += 0.1 * np.cos(x[0] * x[1]) * x[1]
    g[1] += 0.1 * np.cos(x[0] * x[1]) * x[0]
    return f, g


X0 = np.array([1.5, 0.2, -0.7, 0.0, 1.0])
LR = list(0.2 * 0.97 ** np.arange(25))
CASES = [("adam", {}), ("adam", {"beta1": 0.8, "beta2": 0.99, "epsilon": 1e-6}), ("rmsprop", {}),
         ("rmsprop", {"decay": 0.5, "memory": 0.8}), ("adagrad", {}), ("adadelta", {}), ("adadelta", {"rho": 0.8, "epsilon": 1e-4})]

if __name__ == "__main__":
    rh.load_reference()
    from spdepy.optim import Optimize
    from spdepy.optim.sgd import SGD
    os.makedirs(OUT, exist_ok=True)
    out = {"x0": X0, "lr": np.array(LR), "A": A, "C": C}
    for i, (step, hp) in enumerate(CASES):
        for fix in (None, [1, 3]):
            opt = Optimize(fun)
            with contextlib.redirect_stdout(io.StringIO()):
                res = opt.fit(x0=X0.copy(), lr=LR, stepType=step, pol=5, fix=fix, **hp)
            key = "%d_%s_%s" % (i, step, "fix" if fix else "free")
            out[key + "_hist"] = np.array(opt.histX)
            out[key + "_x"] = res["x"]
            out[key + "_f"] = res["fun"]
            out[key + "_jac"] = res["jac"]
    # the reference's SGD has no printInit, so its own fit() cannot drive it: replay fit's loop body on the class
    sgd = SGD()
    x = X0.copy()
    hist = []
    for k in range(len(LR)):
        f, g = fun(x)
        x = sgd(x, f, g, LR[k])
        hist.append(x)
    out["sgd_hist"] = np.array(hist)
    np.savez_compressed(os.path.join(OUT, "trajectories.npz"), **out)
    print("wrote", len(out), "arrays")
