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
        self._Q = None
        self._Qdev = None
        self.Q_fac = None
        self.mvar = None
        self.mu = np.zeros(int(np.prod(self.grid.shape)))
        self.useCov = False
        self.sigmas = np.log(np.array([0.01, 140]))

    # ``Q``: the precision as SciPy CSC, as in the reference (``model.py:123,132,152``).  After ``update`` /
    # ``setModel(useCov=True)`` the matrix lives on the device and is exported on first read.
    @property
    def Q(self):
        return self.getQ()

    @Q.setter
    def Q(self, value):
        self._Q = value

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
        self.optim = Optimize(self._objective)

    def _objective(self, par):
        """``mod.logLike`` as the optimiser calls it (``optim/__init__.py:44``).  Meshes evaluated by the streamed path
        (posterior factor beyond device memory) have no Hutchinson mode -- it needs two resident 3-D factors -- so
        they are fitted with the exact Takahashi gradient."""
        if getattr(self.mod, "timed", False) and self.mod.engine.use_streamed():
            return self.mod.logLike(par, exact_grad=True)
        return self.mod.logLike(par)

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
        """``model.py:129-153``.  With ``useCov`` the latent vector gains 1-2 regression coefficients: ``S`` gets
        dense columns (``grids.addCov/addInt``) and ``Q`` a diagonal block ``diag(exp(sigmas))``; after
        ``update`` the precision is *bordered*, ``[[Q11, B], [B^T, C]]`` with ``B`` n x k dense.  The sparse block
        stays on the device in the slot layout and keeps the supernodal factor; the border is carried by a
        k x k Schur complement (:class:`_Border`)."""
        self.useCov = useCov if useCov is not None else self.useCov
        if self.mod._state is None:
            self.mod.setQ()
        self._Qdev = self.mod._state["Q"].clone()
        self.tau = np.exp(self.mod.tau)
        self._border = None
        if useCov:      # (the argument, not self.useCov: model.py:131)
            self.sigmas = sigmas if sigmas is not None else self.sigmas
            n = self.mod.engine.n
            if hasattr(self.sigmas, "__len__"):
                self.grid.addCov(mu, scale=scale)
                diag = np.exp(np.asarray(self.sigmas, dtype="float64"))
                if scale:
                    self.mu = np.zeros(n + diag.size)
                    self.mu[-2:] = [0, np.max(mu)]
                else:
                    self.mu = np.zeros(n + diag.size) if mu is None else self.grid.getS().T @ mu
            else:
                self.grid.addInt()
                diag = np.array([np.exp(self.sigmas)], dtype="float64")
                S = self.grid.getS()
                self.mu = np.zeros(S.shape[1])
                self.mu[:-1] = S[:, :-1].T @ mu
            self._border = _Border(n, diag)
            self.Q = None
            return
        self.Q = self.mod.Q.copy().tocsc()
        self.mu = np.zeros(self.Q.shape[0]) if mu is None else self.grid.getS().T @ mu

    # ------------------------------------------------------------------ model.py:73-87
    def sample(self, n=1, simple=False, seed=None, cols=None) -> np.ndarray:
        """``cols`` (a slice, optional) restricts the work to those columns of the same ``n``-column draw: the
        column-block shard of one rank (``spdepy_b200.parallel.sample_sharded``)."""
        if seed is None:
            seed = np.random.randint(100)
        eng = self.mod.engine
        if self.useCov:
            return self._sample_bordered(n, simple, seed, cols)
        N = eng.n
        z = np.random.default_rng(seed).normal(size=N * n).reshape(N, n)
        if cols is not None:
            z = np.ascontiguousarray(z[:, cols])
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
        if self.useCov:
            return self._update_bordered(y, idx, tau)
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
        if self._Q is None and self._Qdev is not None:
            Q11 = self.mod.engine.to_scipy(self._Qdev)
            if self.useCov and self._border is not None:
                from scipy import sparse
                B = sparse.csc_matrix(self._border.B.cpu().numpy())
                Q11 = sparse.bmat([[Q11, B], [B.T, sparse.csc_matrix(self._border.C)]]).tocsc()
                Q11.eliminate_zeros()
            self._Q = Q11
        return self._Q

    # ------------------------------------------------------------------ bordered precision (useCov=True)
    def _cov_columns(self, idx=None):
        """Dense regression columns of ``S`` (rows = observations ``idx``) and the mesh nodes of those rows."""
        S = self.grid.getS(idx)
        n = self.mod.engine.n
        X = np.asarray(S[:, n:].todense(), dtype="float64")
        nodes = np.asarray(self.grid.obs_nodes(idx), dtype=np.int64)
        return X, nodes

    def _sample_bordered(self, n, simple, seed, cols=None):
        eng, bd = self.mod.engine, self._border
        N, k = eng.n, bd.k
        z = np.random.default_rng(seed).normal(size=(N + k) * n).reshape(N + k, n)
        if cols is not None:
            z = np.ascontiguousarray(z[:, cols])
        self.Q_fac = eng.factorize(0, self._Qdev)
        Y, L22 = bd.factor(eng)
        # [[L11, 0], [Y^T, L22]]^T u = z  ->  u2 = L22^-T z2,  u1 = L11^-T (z1 - Y u2);  x = [P^T u1; u2]
        u2 = np.linalg.solve(L22.T, z[N:])
        rhs = to_dev(z[:N]) - Y @ to_dev(u2)
        x1 = eng.solve(0, rhs.contiguous(), 10)
        x1 += to_dev(self.mu[:N])[:, None]
        X, nodes = self._cov_columns()
        data = x1[torch.as_tensor(nodes, device=x1.device)].cpu().numpy() + X @ (u2 + self.mu[N:, None])
        if not simple:
            data += z[nodes] * 1 / np.sqrt(self.tau)
        return data

    def _update_bordered(self, y, idx, tau):
        eng, bd = self.mod.engine, self._border
        N = eng.n
        X, nodes = self._cov_columns(idx)
        cnt = to_dev(np.bincount(nodes, minlength=N).astype(np.float64))
        eng.add_diag(self._Qdev, cnt, tau)                      # Q11 + tau S1^T S1
        nodes_d = to_dev(nodes, torch.int64)
        bd.B.index_add_(0, nodes_d, to_dev(X) * tau)            # B + tau S1^T X
        bd.C = bd.C + tau * X.T @ X
        self.Q = None
        self.Q_fac = eng.factorize(0, self._Qdev)
        resid = np.asarray(y, dtype="float64") - (self.mu[nodes] + X @ self.mu[N:])
        b1 = eng.scatter_obs(to_dev(resid.reshape(-1, 1)), nodes_d, 1.0)
        b2 = X.T @ resid
        x1, x2 = bd.solve(eng, b1, b2)
        self.mu = self.mu + np.concatenate([x1.cpu().numpy()[:, 0], x2]) * tau

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
        mvar = Z[nd * eng.n:(nd + 1) * eng.n]
        if self.useCov:
            # Marginal variances of the bordered latent vector [x; beta] (the reference feeds ``S Q S^T`` with the
            # shape of Q to R/INLA here, ``model.py:108-116``, which is not a precision matrix; this is the inverse of
            # the precision the rest of the facade uses): with W = Q11^-1 B and the Schur complement
            # Sc = C - B^T W,  diag(Z11) = diag(Q11^-1) + rowsum((W Sc^-1) * W),  Z22 = Sc^-1.
            bd = self._border
            W = eng.solve(0, bd.B.clone())
            Sci = np.linalg.inv(bd.C - (bd.B.T @ W).cpu().numpy())
            mvar = mvar + ((W @ to_dev(Sci)) * W).sum(dim=1)
            self.mvar = np.concatenate([mvar.cpu().numpy(), np.diag(Sci)])
            return self.mvar
        self.mvar = mvar.cpu().numpy()
        return self.mvar


class _Border:
    """Dense border of the precision ``[[Q11, B], [B^T, C]]`` (``useCov=True``): ``B`` (n x k, device), ``C`` (k x k,
    host).  Everything goes through the sparse factor of ``Q11`` (engine store 0) and the k x k Schur complement
    ``C - B^T Q11^-1 B``; with the border ordered last, the Cholesky factor of the bordered matrix is
    ``[[L11, 0], [Y^T, L22]]`` with ``Y = L11^-1 P B`` and ``L22 L22^T = C - Y^T Y``."""

    def __init__(self, n, diag):
        self.k = int(diag.size)
        self.B = torch.zeros(n, self.k, dtype=F64, device=torch.device("cuda", torch.cuda.current_device()))
        self.C = np.diag(diag).astype("float64")

    def factor(self, eng):
        Y = eng.solve(0, self.B.clone(), 5)                     # L11^-1 P B (rows in the factor's ordering)
        S = self.C - (Y.T @ Y).cpu().numpy()
        return Y, np.linalg.cholesky(S)

    def solve(self, eng, b1, b2):
        v = eng.solve(0, b1.clone())                            # Q11^-1 b1
        W = eng.solve(0, self.B.clone())                        # Q11^-1 B
        S = self.C - (self.B.T @ W).cpu().numpy()
        x2 = np.linalg.solve(S, b2 - (self.B.T @ v).cpu().numpy()[:, 0])
        x1 = v - W @ to_dev(x2.reshape(-1, 1))
        return x1, x2
