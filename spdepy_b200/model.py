"""``Model`` facade with the reference's surface (``model.py:7-163``): ``fit / sample / update /
setModel / setQ / getPars / qinv``.  Sampling, conditioning and marginal variances run on the
device through the engine; ``qinv(simple=False)`` uses the Takahashi selected inverse instead of
the reference's R/INLA subprocess (``rqinv.R``)."""
from __future__ import annotations

import numpy as np
import torch

from .engine import F64, Engine, to_dev
from .optim import Optimize
from .spdes import spde_init


class Model:
    def __init__(self, spde=None, grid=None, parameters=None, ani=True, ha=True, bc=3, mod0=None) -> None:
        self.grid = grid
        self.mod = None
        if spde is not None:
            self.model(spde=spde, grid=grid, parameters=parameters, ani=ani, ha=ha, bc=bc, mod0=mod0)
        self.Q = None
        self._Qdev = None
        self.Q_fac = None
        self.mvar = None
        self.mu = np.zeros(int(np.prod(self.grid.shape)))
        self.useCov = False
        self.sigmas = np.log(np.array([0.01, 140]))

    def setQ(self, par=None) -> None:
        self.mod.setQ(par=par)

    def model(self, spde=None, grid=None, parameters=None, ani=True, ha=True, bc=3, mod0=None) -> None:
        assert spde is not None and grid is not None
        if grid.type == "gridST" and mod0 is None:
            from .grids import grid as make_grid
            mod0 = spde_init(model="whittle-matern", grid=make_grid(x=grid.x, y=grid.y, extend=grid.Ne or None),
                             ani=ani, ha=ha, bc=bc)
        self.mod = spde_init(model=spde, grid=grid, parameters=parameters, ani=ani, ha=ha, bc=bc, mod0=mod0)
        self.spde_type = self.mod.type
        self.optim = Optimize(self.mod.logLike)

    def fit(self, data, **kwargs):
        assert self.mod is not None
        x0 = self.mod.initFit(data, **kwargs)
        if kwargs.get("x0") is None:
            kwargs["x0"] = x0
        if kwargs.get("verbose") is not None:
            kwargs["print"] = self.mod.print
        for k in ("idx", "fitQ0"):
            kwargs.pop(k, None)
        res = self.optim.fit(**kwargs)
        self.setQ(par=res["x"])
        return res

    def getPars(self, onlySelf=True) -> np.ndarray:
        return self.mod.getPars(onlySelf=onlySelf)

    # ------------------------------------------------------------------ model.py:129-153
    def setModel(self, mu=None, sigmas=None, useCov=None, scale=True):
        if useCov:
            raise NotImplementedError("regression columns in S (useCov=True) border the precision matrix; "
                                      "that case is SURVEY.md section 8f #1, not part of this round")
        if self.mod._state is None:
            self.mod.setQ()
        self.useCov = False
        self._Qdev = self.mod._state["Q"].clone()
        self.Q = self.mod.Q.copy().tocsc()
        self.mu = np.zeros(self.Q.shape[0]) if mu is None else self.grid.getS().T @ mu
        self.tau = np.exp(self.mod.tau)

    # ------------------------------------------------------------------ model.py:73-87
    def sample(self, n=1, simple=False, seed=None) -> np.ndarray:
        if seed is None:
            seed = np.random.randint(100)
        eng = self.mod.engine
        N = eng.n
        z = np.random.default_rng(seed).normal(size=N * n).reshape(N, n)
        self.Q_fac = eng.factorize(0, self._Qdev)
        x = eng.solve(0, to_dev(z), 10)                         # P^T L^-T z
        x += to_dev(self.mu)[:, None]
        nodes = torch.as_tensor(self.grid.obs_nodes(), device=x.device)
        data = x[nodes].cpu().numpy()                           # S @ (...)
        if not simple:
            data += z[self.grid.obs_nodes()] * 1 / np.sqrt(self.tau)
        return data

    # ------------------------------------------------------------------ model.py:120-127
    def update(self, y, idx, tau=None):
        if tau is None:
            tau = self.tau
        eng = self.mod.engine
        nodes = np.asarray(self.grid.obs_nodes(idx), dtype=np.int64)
        cnt = to_dev(np.bincount(nodes, minlength=eng.n).astype(np.float64))
        eng.add_diag(self._Qdev, cnt, tau)                      # Q + tau S^T S
        self.Q = None                                           # exported lazily, see getQ()
        self.Q_fac = eng.factorize(0, self._Qdev)
        resid = np.asarray(y, dtype="float64") - self.mu[nodes]
        b = eng.scatter_obs(to_dev(resid.reshape(-1, 1)), to_dev(nodes, torch.int64), 1.0)
        tmp = eng.solve(0, b).cpu().numpy()[:, 0] * tau
        self.mu = self.mu + tmp

    def getQ(self):
        if self.Q is None and self._Qdev is not None:
            self.Q = self.mod.engine.to_scipy(self._Qdev)
        return self.Q

    # ------------------------------------------------------------------ model.py:89-118
    def qinv(self, simple=False):
        if simple:
            z = self.sample(n=1000, simple=True)
            return z.var(axis=1)
        if self._Qdev is None:
            self.setModel()
        eng = self.mod.engine
        eng.factorize(0, self._Qdev)
        Z = eng.selinv(0)
        nd = eng.nslots // 2
        self.mvar = Z[nd * eng.n:(nd + 1) * eng.n].cpu().numpy()
        return self.mvar
