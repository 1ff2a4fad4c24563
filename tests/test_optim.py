"""The host optimiser driver (spdepy_b200/optim, the caller side of logLike: SURVEY.md section 8f #2) against
trajectories recorded from the UNMODIFIED reference optimiser (oracle/make_golden_optim.py ->
tests/golden/optim/trajectories.npz): every iterate of every update rule bit for bit, with and without fixed
coordinates, Polyak average and result dictionary included."""
import os

import numpy as np
import pytest

from spdepy_b200.optim import Optimize

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "optim", "trajectories.npz"))
A, C, X0, LR = G["A"], G["C"], G["x0"], list(G["lr"])
CASES = [("adam", {}), ("adam", {"beta1": 0.8, "beta2": 0.99, "epsilon": 1e-6}), ("rmsprop", {}),
         ("rmsprop", {"decay": 0.5, "memory": 0.8}), ("adagrad", {}), ("adadelta", {}), ("adadelta", {"rho": 0.8, "epsilon": 1e-4})]


def fun(x):
    d = x - C
    f = float((A * d ** 2).sum() + 0.1 * np.sin(x[0] * x[1]))
    g = 2 * A * d
    g[0] += 0.1 * np.cos(x[0] * x[1]) * x[1]
    g[1] += 0.1 * np.cos(x[0] * x[1]) * x[0]
    return f, g


@pytest.mark.parametrize("i", range(len(CASES)))
@pytest.mark.parametrize("fix", [None, [1, 3]])
def test_iterates_match_the_reference(i, fix):
    step, hp = CASES[i]
    key = "%d_%s_%s" % (i, step, "fix" if fix else "free")
    opt = Optimize(fun)
    res = opt.fit(x0=X0.copy(), lr=LR, stepType=step, pol=5, fix=fix, **hp)
    assert np.array_equal(np.array(opt.histX), G[key + "_hist"])
    assert np.array_equal(res["x"], G[key + "_x"]) and res["fun"] == float(G[key + "_f"])
    assert np.array_equal(res["jac"], G[key + "_jac"]) and res["method"] == step
    if fix:
        assert np.array_equal(np.array(opt.histX)[:, fix], np.repeat(X0[None, fix], len(LR), axis=0))
    assert len(opt.histF) == len(opt.histJac) == len(LR)


def test_momentum_sgd_and_extensions(tmp_path):
    opt = Optimize(fun)
    opt.fit(x0=X0.copy(), lr=LR, stepType="sgd")
    assert np.array_equal(np.array(opt.histX), G["sgd_hist"])
    # scalar learning rate (the reference raises), eps alias, saved result, sticky step type
    res = opt.fit(x0=X0.copy(), lr=0.1, max_steps=7, stepType="adam", eps=1e-6, end=str(tmp_path / "fit"))
    ref = Optimize(fun).fit(x0=X0.copy(), lr=[0.1] * 7, stepType="adam", epsilon=1e-6)
    assert len(opt.histX) == 7 and np.array_equal(res["x"], ref["x"])
    assert np.array_equal(np.load(str(tmp_path / "fit.npy")), res["x"])
    with pytest.raises(ValueError):
        opt.fit(x0=X0.copy(), lr=LR, stepType="newton")


def test_adam_first_step_convention():
    """optim/adam.py:14-16: moments start at g and g^2, so step one is lr * sqrt(1 - beta2) / (1 - beta1) * sign(g)."""
    opt = Optimize(lambda x: (0.0, np.array([3.0, -0.5])))
    opt.fit(x0=np.zeros(2), lr=[0.1], stepType="adam", epsilon=0.0)
    assert np.allclose(opt.histX[0], -0.1 * np.sqrt(1 - 0.999) / (1 - 0.9) * np.array([1.0, -1.0]), rtol=1e-12)
