"""Sequential stochastic-gradient driver with the reference's call contract
(``optim/__init__.py:26-62``: ``fun(x) -> (f, jac)``, per-step learning rates, ``fix=`` indices,
Polyak average of the last ``pol`` iterates).  Pure host scalar math around the device hot path."""
import numpy as np


class Optimize:
    def __init__(self, fun):
        self.fun = fun
        self.histX, self.histF, self.histJac = [], [], []

    def fit(self, x0=None, lr=None, max_steps=None, fix=None, pol=10, stepType="adam", verbose=False, end=None,
            beta1=0.9, beta2=0.999, eps=1e-8, **kwargs):
        x = np.array(x0, dtype="float64")
        if lr is None:
            lr = [0.1] * (max_steps or 100)
        lr = np.atleast_1d(np.asarray(lr, dtype="float64"))
        if max_steps is not None and lr.size == 1:
            lr = np.repeat(lr, max_steps)
        fix = [] if fix is None else list(fix)
        m, v = np.zeros_like(x), np.zeros_like(x)
        prt = kwargs.get("print")
        self.histX, self.histF, self.histJac = [], [], []
        for k, step in enumerate(lr):
            f, jac = self.fun(x)
            jac = np.array(jac, dtype="float64")
            jac[fix] = 0.0
            if stepType == "sgd":
                upd = jac
            else:                                   # Adam (optim/adam.py:12-25)
                m = beta1 * m + (1 - beta1) * jac
                v = beta2 * v + (1 - beta2) * jac ** 2
                upd = (m / (1 - beta1 ** (k + 1))) / (np.sqrt(v / (1 - beta2 ** (k + 1))) + eps)
            x = x - step * upd
            self.histX.append(x.copy()); self.histF.append(f); self.histJac.append(jac)
            if verbose and prt is not None:
                print("# %4d | f = %2.4f" % (k, f), prt(x))
        tail = np.array(self.histX[-pol:]) if self.histX else x[None, :]
        xbar = tail.mean(axis=0)
        if end is not None:
            np.save(end, xbar)
        return {"x": xbar, "fun": self.histF[-1] if self.histF else None, "jac": self.histJac[-1] if self.histJac else None}
