"""Sequential stochastic-gradient driver with the reference's call contract (``optim/__init__.py:26-62``):
``fun(x) -> (f, jac)``, one learning rate per step, ``fix=`` indices held at their start values, Polyak average of
the last ``pol`` iterates, result dictionary ``{method, x, fun, jac}``.  Pure host scalar math around the device hot
path; the five update rules reproduce the reference's iterates bit for bit (``tests/test_optim.py`` against
trajectories recorded from the unmodified reference, ``oracle/make_golden_optim.py``), including their first-step
conventions:

* ``adam`` (``optim/adam.py:12-25``): the moments START at ``g`` and ``g**2`` (not at zero) and are still divided by
  ``1 - beta**t``, so the first step is ``lr * sqrt(1 - beta2) / (1 - beta1)`` per coordinate;
* ``rmsprop`` (``optim/rmsprop.py:10-18``): running mean of ``g**2`` started at ``g**2``, exact zeros replaced by 1,
  optional momentum ``decay`` (0 by default);
* ``adagrad`` (``optim/ada_grad.py:8-14``), ``adadelta`` (``optim/ada_delta.py:11-20``, ignores the learning rate);
* ``sgd`` (``optim/sgd.py:8-10``): heavy-ball momentum 0.9.  (The reference's ``SGD`` lacks ``printInit`` and cannot
  be driven through its own ``fit``; here it can.)

Extensions that do not change the reference's behaviour: a scalar ``lr`` is repeated ``max_steps`` times (the
reference raises there), ``eps=`` is accepted for ``epsilon=``.
"""
import numpy as np


def _adam(st, x, g, lr, hp):
    b1, b2, eps = hp.get("beta1", 0.9), hp.get("beta2", 0.999), hp.get("epsilon", 1e-8)
    st["t"] = st.get("t", 0) + 1
    if "m" not in st:
        st["m"], st["v"] = g, g ** 2
    else:
        st["m"] = b1 * st["m"] + (1 - b1) * g
        st["v"] = b2 * st["v"] + (1 - b2) * g ** 2
    mhat = st["m"] / (1 - b1 ** st["t"])
    vhat = st["v"] / (1 - b2 ** st["t"])
    return x - lr * mhat / (np.sqrt(vhat) + eps)


def _rmsprop(st, x, g, lr, hp):
    decay, memory = hp.get("decay", 0.0), hp.get("memory", 0.9)
    if "g2" not in st:
        st["g2"] = g ** 2
    else:
        st["g2"] = memory * st["g2"] + (1 - memory) * g ** 2
    st["g2"][st["g2"] == 0] = 1
    st["dx"] = decay * st.get("dx", 0) - lr * g / np.sqrt(st["g2"])
    return x + st["dx"]


def _adagrad(st, x, g, lr, hp):
    if "g2" not in st:
        st["g2"] = g ** 2
    else:
        st["g2"] += g ** 2
    return x + (-lr * g / (np.sqrt(st["g2"]) + 1e-8))        # the rule's own epsilon argument, not the setting (ada_grad.py:8)


def _adadelta(st, x, g, lr, hp):
    rho, eps = hp.get("rho", 0.9), hp.get("epsilon", 1e-8)
    if "eg2" not in st:
        st["eg2"] = (1 - rho) * g ** 2
        dx = -np.sqrt(eps) / np.sqrt(st["eg2"] + eps) * g
        st["edx2"] = (1 - rho) * dx ** 2
    else:
        st["eg2"] = rho * st["eg2"] + (1 - rho) * g ** 2
        dx = -np.sqrt(st["edx2"] + eps) / np.sqrt(st["eg2"] + eps) * g
        st["edx2"] = rho * st["edx2"] + (1 - rho) * dx ** 2
    return x + dx


def _sgd(st, x, g, lr, hp):
    st["dx"] = hp.get("decay", 0.9) * st.get("dx", 0) - lr * g
    return x + st["dx"]


RULES = {"adam": _adam, "rmsprop": _rmsprop, "adagrad": _adagrad, "adadelta": _adadelta, "sgd": _sgd}


class Optimize:
    def __init__(self, fun=None):
        self.fun = fun
        self.stepType = "adam"
        self.pol = 10
        self.histX, self.histF, self.histJac = [], [], []
        self.x = self.f = self.jac = None

    def fit(self, x0=None, lr=None, max_steps=None, fix=None, pol=None, stepType=None, verbose=False, end=None,
            fun=None, **kwargs):
        if fun is not None:
            self.fun = fun
        assert self.fun is not None
        if stepType is not None:
            self.stepType = stepType
        if self.stepType not in RULES:
            raise ValueError("Step type not recognized")
        rule = RULES[self.stepType]
        if pol is not None:
            self.pol = pol
        if x0 is not None:
            self.x = x0
        x = np.array(self.x, dtype="float64")
        if lr is None:
            lr = 0.1
        if not hasattr(lr, "__len__"):
            lr = [lr] * (max_steps or 100)
        nsteps = len(lr) if max_steps is None else min(int(max_steps), len(lr))
        hp = {k: v for k, v in kwargs.items() if v is not None and k in ("beta1", "beta2", "epsilon", "decay", "memory", "rho")}
        if kwargs.get("eps") is not None:
            hp.setdefault("epsilon", kwargs["eps"])
        prt = kwargs.get("print")
        state = {}
        self.histX, self.histF, self.histJac = [], [], []
        for k in range(nsteps):
            self.f, self.jac = self.fun(x)
            jac = np.asarray(self.jac, dtype="float64")
            if verbose:
                print("# %3.0d" % k, "| fun = %2.4f" % (self.f), prt(x) if prt is not None else "")
            self.histF.append(self.f)
            self.histJac.append(self.jac)
            keep = x[fix] if fix is not None else None
            x = rule(state, x, jac, lr[k], hp)
            if fix is not None:
                x[fix] = keep
            self.histX.append(x)
        self.x = np.array(self.histX).T[:, -self.pol:].mean(axis=1) if self.histX else x
        res = {"method": self.stepType, "x": self.x, "fun": self.f, "jac": self.jac}
        if end is not None:
            np.save(end + ".npy", res["x"])
        return res
