"""Separable space-time models ``Q = Qt (x) Qs`` (``seperable_spatial_temporal2D.py``, ``..._ha2D.py``,
``..._idiffusion2D.py``): tridiagonal precision in time, spatially varying Whittle-Matern precision in space.

Reference layouts (Np hard-coded 9):

* anisotropic (``seperable_spatial_temporal2D.py:30-38``): ``par = [kappa x9, gamma x9, vx x9, vy x9, log rho, log tau]``,
  ``Qt`` the AR(1) precision in ``rho`` (``:186-211``);
* half-angle (``seperable_spatial_temporal_ha2D.py:26-35``): ``par = [kappa x9, gamma x9, vx x9, vy x9, a, log sigma, log tau]``;
* isotropic (``seperable_spatial_temporal_idiffusion2D.py:30-36``): ``par = [kappa x9, gamma x9, a, log sigma, log tau]``;
  for both, ``Qt = sigma tridiag(-a, 1 + a^2, -a)`` with ``sigma`` alone in the two corners, its derivative list is
  ``[kron(dQt/da, Qs), Q]`` (``..._ha2D.py:133-136``), and the reference fills ``Qt`` with a hard-coded ``range(10)``
  (``..._ha2D.py:205,216``): the classes exist for ``T = 10`` only and say so here instead of returning a singular matrix.

``Q`` has 75 entries per row (5x5 blocks to t-1, t and t+1), the third slot layout of the C ABI
(``SPDE_PATTERN_KRON``); the supernodal plan, the solves, the Takahashi pass and the reductions are the same kernels
as for the advection-diffusion family.  The spatial factor ``Qs = As^T Dv^-1 As`` (``:78-80``) is assembled by the
var-Whittle-Matern configuration of :class:`SPDE2D`, which also owns the chain rule of every spatial parameter: the
weights on ``Q`` are collapsed over time with ``Qt`` (``spde_kron_reduce``) and handed to its adjoint pass.

The prior never needs a 3-D factorisation: ``logdet Q = Ns logdet Qt + T logdet Qs`` and
``tr(Q^-1 dQ_i) = T tr(Qs^-1 dQs_i)`` (spatial parameters), ``Ns tr(Qt^-1 dQt_j)`` (temporal parameters).
"""
from __future__ import annotations

import numpy as np
import torch

from ..engine import F64, Engine, to_dev, to_host
from .base import LazyDQ, SPDE2D


class _SpatialFactor(SPDE2D):
    """Qs of the separable model: spline kappa, spline anisotropic H, no advection (``:70-80``)."""
    name = "seperable-spatial-temporal-ani-2D[Qs]"
    timed = False
    kvar = True
    Hkind = "aniso"
    Hvar = True
    default_own = ([-1] * 9, [-1] * 9, [0.1] * 9, [0.1] * 9)


class _SpatialFactorHa(_SpatialFactor):
    """Half-angle H = gamma (cosh|v| I + sinh|v|/|v| [[vx, vy], [vy, -vx]]) (``seperable_spatial_temporal_ha2D.py:82-89``)."""
    name = "seperable-spatial-temporal-ha-2D[Qs]"
    Hkind = "ha"


class _SpatialFactorIso(_SpatialFactor):
    """Isotropic H = gamma I (``seperable_spatial_temporal_idiffusion2D.py:74``)."""
    name = "seperable-spatial-temporal-2D[Qs]"
    Hkind = "iso"
    default_own = ([-1] * 9, [-1] * 9)


def qt_coeffs(rho: float, diff: int = 0):
    """(d0, d1, e) of the tridiagonal ``makeQt`` (``:186-211``): diagonal (d0, d1, ..., d1, d0), off-diagonal e;
    ``diff=1`` is the derivative with respect to log rho."""
    if diff == 1:
        return (2 * rho ** 2 / (1 - rho ** 2) ** 2, 4 * rho ** 2 / (1 - rho ** 2) ** 2,
                -rho * (1 + rho ** 2) / (1 - rho ** 2) ** 2)
    return 1 / (1 - rho ** 2), (1 + rho ** 2) / (1 - rho ** 2), -rho / (1 - rho ** 2)


def qt_coeffs_a(a: float, sigma: float, diff: int = 0):
    """(d0, d1, e) of ``makeQt(a, sigma)`` of the ha / idiffusion classes (``seperable_spatial_temporal_ha2D.py:202-226``);
    ``diff=1`` is the derivative with respect to ``a``."""
    if diff == 1:
        return 0.0, 2 * a * sigma, -sigma
    return sigma, (1 + a ** 2) * sigma, -a * sigma


def _qt_dense(T, c):
    d0, d1, e = c
    Q = np.zeros((T, T))
    for i in range(T):
        Q[i, i] = d0 if i in (0, T - 1) else d1
        if i > 0:
            Q[i, i - 1] = e
        if i < T - 1:
            Q[i, i + 1] = e
    return Q


class SeperableSpatialTemporal2D:
    """Drop-in for ``spdepy.spdes.seperable_spatial_temporal2D.SeperableSpatialTemporal2D`` (same spelling)."""
    timed = True
    collapse_prior = True
    _spatial_cls = _SpatialFactor
    _type = "seperable-spatial-temporal-ani-2D-bc%d"
    _nsp = 36                 # spatial parameters; the temporal ones follow, log tau is last

    def __init__(self, grid, par=None, bc=3) -> None:
        self.grid = grid
        self.type = self._type % bc
        self.Q = None
        self.Q_fac = None
        self.data = None
        self.r = None
        self.S = None
        self.bc = bc
        self.mod0 = None
        self._state = None
        self._obs = None
        M, N, T = grid.shape[0], grid.shape[1], grid.T
        if T < 2:
            raise ValueError("the separable space-time model needs at least two time steps")
        self._check_T(T)
        self.engine = Engine.get(M, N, T, bc, pat=1)
        self.spatial = self._spatial_cls(grid, bc=bc)
        if par is None:
            self.setPars(self._default_par())
        else:
            self.setQ(par=par)

    def _check_T(self, T):
        pass

    # ------------------------------------------------------------------ parameters (:28-47)
    @staticmethod
    def _default_par():
        return np.hstack([[-1] * 9, [-1] * 9, [0.1] * 9, [0.1] * 9, -1, np.log(100)]).astype("float64")

    def getPars(self, *args, **kwargs) -> np.ndarray:
        return np.hstack([self.kappa, self.gamma, self.vx, self.vy, self.rho, self.tau]).astype("float64")

    def setPars(self, par) -> None:
        par = np.array(par, dtype="float64")
        self.kappa, self.gamma = par[0:9], par[9:18]
        self.vx, self.vy = par[18:27], par[27:36]
        self.rho, self.tau = par[36], par[37]

    def _qt_list(self, par):
        """Coefficients (d0, d1, e) of ``Qt`` and of its derivative for every temporal parameter, in ``par`` order."""
        rho = float(np.exp(par[36]))
        return qt_coeffs(rho), [qt_coeffs(rho, 1)]            # d / d log rho

    def initFit(self, data, **kwargs):
        data = np.asarray(data, dtype="float64")
        assert data.shape[0] <= self.grid.n
        self.data = data
        self.r = data.shape[1] if data.ndim == 2 else 1
        idx = kwargs.get("idx")
        self.S = self.grid.getS(idxs=idx)
        self._set_obs(self.grid.obs_nodes(idx))
        return self._default_par()

    def _set_obs(self, nodes):
        nodes = np.asarray(nodes, dtype=np.int64)
        cnt = np.bincount(nodes, minlength=self.engine.n).astype(np.float64)
        self._obs = {"nodes": to_dev(nodes, torch.int64), "cnt": to_dev(cnt), "nobs": int(nodes.size)}

    def setQ(self, par=None, S=None):
        if par is None:
            par = self.getPars()
        else:
            self.setPars(par)
        if S is not None:
            self.S = S
        self.Q, self.Q_fac = self.makeQ(par=np.asarray(par, dtype="float64"), grad=False)
        self.S = self.grid.getS()
        self._set_obs(self.grid.obs_nodes())

    def print(self, par):
        return ("| κ = %2.2f" % (np.exp(par[0:9]).mean()) + ", γ = %2.2f" % (np.exp(par[9:18]).mean())
                + ", vx = %2.2f" % ((par[18:27]).mean()) + ", vy = %2.2f" % ((par[27:36]).mean())
                + ", ρ = %2.2f" % (np.exp(par[36])) + ", τ = %2.2f" % (np.exp(par[37])))

    def setClib(self) -> None:
        return None

    def Ah(self, Hs):
        return self.spatial.Ah(Hs)

    def makeQt(self, rho, T=10, diff=0):
        from scipy import sparse
        return sparse.csc_matrix(_qt_dense(T, qt_coeffs(rho, diff)))

    # ------------------------------------------------------------------ assembly
    def _assemble(self, par):
        par = np.asarray(par, dtype="float64")
        sp_par = np.hstack([par[:self._nsp], par[-1]])
        st_s = self.spatial._assemble(sp_par)                 # kappa, A9 = Dv Dk - Ah(Hs), Qs = As^T Dv^-1 As
        qt, dqt = self._qt_list(par)
        Q = self.engine.fill_kron(st_s["Q"], *qt)
        return {"par": par, "spatial": st_s, "qt": qt, "dqt": dqt, "Q": Q, "joint": False}

    def makeQ(self, par, grad=True):
        """``(Q, Q_fac)`` or ``(Q, Q_fac, dQ)`` as the reference (``:64-122``); ``dQ`` are lazy operators."""
        st = self._assemble(np.asarray(par, dtype="float64"))
        self._state = st
        fac = self.engine.factorize(0, st["Q"])
        Q = self.engine.to_scipy(st["Q"])
        if not grad:
            return Q, fac
        return Q, fac, self._lazy_dQ(st)

    def _lazy_dQ(self, st):
        ops = []
        for op in self.spatial._lazy_dQ(st["spatial"]):           # kron(Qt, dQs_i)
            ops.append(_KronDQ(self, st, op, None))
        for dqt in st["dqt"]:
            ops.append(_KronDQ(self, st, None, dqt))              # kron(dQt_j, Qs)
        return ops

    # ------------------------------------------------------------------ gradient contraction
    def _grad_from_weights(self, st, W, prior=None):
        """sum(W .* dQ_i) for the own parameters (all but log tau); ``prior = {"c": c, "Zs": Z of Qs}`` adds
        ``c d logdet Q``."""
        eng = self.engine
        Wd = eng.kron_reduce(W, *st["qt"])
        g_t = [Engine.dot(W, eng.fill_kron(st["spatial"]["Q"], *dqt)) for dqt in st["dqt"]]
        if prior is not None:
            c = prior["c"]
            Wd = Wd + (c * eng.T) * prior["Zs"]
            Qt = _qt_dense(eng.T, st["qt"])
            for j, dqt in enumerate(st["dqt"]):
                g_t[j] += c * eng.Ns * float(np.trace(np.linalg.solve(Qt, _qt_dense(eng.T, dqt))))
        out = list(self.spatial._grad_from_weights(st["spatial"], Wd))
        out.extend(g_t)
        return out

    def _prior_collapsed(self, st, want_grad):
        e2 = self.spatial.engine
        e2.factorize(1, st["spatial"]["Q"])
        ldS = e2.logdet(1)
        sign, ldT = np.linalg.slogdet(_qt_dense(self.engine.T, st["qt"]))
        out = {"logdet": self.engine.Ns * ldT + self.engine.T * ldS}
        if want_grad:
            out["Zs"] = e2.selinv(1)
        return out

    # ------------------------------------------------------------------ likelihood (:125-170)
    def logLike(self, par, nh1=100, grad=True, probes=None, exact_grad=False):
        eng = self.engine
        par = np.asarray(par, dtype="float64")
        if self._obs is None or self.data is None:
            raise RuntimeError("call initFit(data, idx=...) first")
        r, nobs = self.r, self._obs["nobs"]
        obs, cnt = self._obs["nodes"], self._obs["cnt"]
        data = to_dev(self.data.reshape(nobs, r))
        tau = float(np.exp(par[-1]))
        st = self._assemble(par)
        self._state = st
        Q = st["Q"]
        collapsed = self.collapse_prior and (exact_grad or not grad)
        if collapsed:
            eng.factorize_async(1, Q, cnt, tau)
            prior = self._prior_collapsed(st, want_grad=grad)
            eng.factor_wait(1)
            ldQ = prior["logdet"]
        else:
            prior = None
            eng.factorize_async(0, Q)
            eng.factorize_async(1, Q, cnt, tau)
            eng.factor_wait(0)
            eng.factor_wait(1)
            ldQ = eng.logdet(0)
        ldQc = eng.logdet(1)
        if grad and exact_grad and collapsed:
            eng.selinv_start(1)
        mu_c = eng.solve(1, eng.scatter_obs(data, obs, tau))
        quad = Engine.dot(mu_c, eng.q_apply(Q, mu_c))
        resid = Engine.residual_ss(data, mu_c, obs)
        like = 1 / 2 * ldQ * r + nobs * r * np.log(tau) / 2 - 1 / 2 * ldQc * r - 1 / 2 * quad - tau / 2 * resid
        self.last = {"mu_c": mu_c, "logdetQ": ldQ, "logdetQc": ldQc, "quad": quad, "resid": resid}
        if not grad:
            return -like / (nobs * r)
        nd = eng.nslots // 2
        if exact_grad and collapsed:
            W = eng.selinv_fetch(1)
            tr_tau = Engine.dot(cnt, W[nd * eng.n:(nd + 1) * eng.n]) * tau
            W *= -0.5 * r
            prior["c"] = 0.5 * r
        elif exact_grad:
            Z, Zc = eng.selinv_pair()
            tr_tau = Engine.dot(cnt, Zc[nd * eng.n:(nd + 1) * eng.n].contiguous()) * tau
            W = (Z - Zc) * (0.5 * r)
            del Z, Zc
        else:
            if probes is None:
                probes = (2 * np.random.randint(1, 3, self.grid.n * nh1) - 3).reshape(self.grid.n, nh1)
            Vp = to_dev(np.asarray(probes, dtype=np.float64))
            nh1 = Vp.shape[1]
            TrQ = eng.solve(0, Vp.clone())
            TrQc = eng.solve(1, Vp.clone())
            a = 0.5 * r / nh1
            W = eng.sddmm(TrQ, Vp, a)
            W = eng.sddmm(TrQc, Vp, -a, W)
            tr_tau = Engine.wdot(TrQc, Vp, cnt) * tau / nh1
        W = eng.sddmm(mu_c, mu_c, -0.5, W)
        g_par = np.zeros(par.size)
        gi = self._grad_from_weights(st, W, prior)
        g_par[:len(gi)] = gi
        g_par[-1] = nobs * r / 2 - 1 / 2 * tr_tau * r - tau / 2 * resid
        return -like / (nobs * r), -g_par / (nobs * r)


class SeperableSpatialTemporalHa2D(SeperableSpatialTemporal2D):
    """Drop-in for ``spdepy.spdes.seperable_spatial_temporal_ha2D.SeperableSpatialTemporalHa2D``: half-angle diffusion,
    ``Qt = sigma tridiag(-a, 1 + a^2, -a)``; ``par = [kappa x9, gamma x9, vx x9, vy x9, a, log sigma, log tau]``."""
    _spatial_cls = _SpatialFactorHa
    _type = "seperable-spatial-temporal-ha-2D-bc%d"

    def _check_T(self, T):
        if T != 10:
            raise ValueError("%s: the reference fills Qt with a hard-coded range(10) (seperable_spatial_temporal_ha2D.py:"
                             "205,216), so the class is defined for T = 10 only (got T = %d)" % (type(self).__name__, T))

    @staticmethod
    def _default_par():
        return np.hstack([[-1] * 9, [-1] * 9, [0.1] * 9, [0.1] * 9, 0.1, 1, np.log(100)]).astype("float64")

    def getPars(self, *args, **kwargs) -> np.ndarray:
        return np.hstack([self.kappa, self.gamma, self.vx, self.vy, self.a, self.sigma, self.tau]).astype("float64")

    def setPars(self, par) -> None:
        par = np.array(par, dtype="float64")
        self.kappa, self.gamma = par[0:9], par[9:18]
        self.vx, self.vy = par[18:27], par[27:36]
        self.a, self.sigma, self.tau = par[36], par[37], par[38]

    def _qt_list(self, par):
        a, sigma = float(par[self._nsp]), float(np.exp(par[self._nsp + 1]))
        qt = qt_coeffs_a(a, sigma)
        return qt, [qt_coeffs_a(a, sigma, 1), qt]             # d / d a;  d / d log sigma = Qt itself

    def print(self, par):
        return ("| κ = %2.2f" % (np.exp(par[0:9]).mean()) + ", γ = %2.2f" % (np.exp(par[9:18]).mean())
                + ", vx = %2.2f" % ((par[18:27]).mean()) + ", vy = %2.2f" % ((par[27:36]).mean())
                + ", a = %2.2f" % (par[36]) + ", σ = %2.2f" % (np.exp(par[37])) + ", τ = %2.2f" % (np.exp(par[38])))

    def transDiff(self, par=None):
        """``seperable_spatial_temporal_ha2D.py:36-45`` (the constant-coefficient form on the first four entries)."""
        par = self.getPars() if par is None else np.asarray(par, dtype="float64")
        aV = np.sqrt(par[2] ** 2 + par[3] ** 2)
        cosh_aV = (np.exp(aV) + np.exp(-aV)) / 2
        sinh_aV = (np.exp(aV) - np.exp(-aV)) / 2
        self.tgamma = np.exp(par[1]) * (cosh_aV - sinh_aV)
        self.tvx = np.sqrt(np.exp(par[0]) * sinh_aV / aV * (par[2] + aV))
        self.tvy = np.sqrt(np.exp(par[0]) * sinh_aV / aV * (-par[2] + aV))

    def makeQt(self, a, sigma, T=10, diff=0):
        from scipy import sparse
        return sparse.csc_matrix(_qt_dense(T, qt_coeffs_a(a, sigma, diff)))


class SeperableSpatialTemporalIDiffusion2D(SeperableSpatialTemporalHa2D):
    """Drop-in for ``spdepy.spdes.seperable_spatial_temporal_idiffusion2D.SeperableSpatialTemporalIDiffusion2D``:
    isotropic diffusion; ``par = [kappa x9, gamma x9, a, log sigma, log tau]``."""
    _spatial_cls = _SpatialFactorIso
    _type = "seperable-spatial-temporal-2D-bc%d"
    _nsp = 18

    @staticmethod
    def _default_par():
        return np.hstack([[-1] * 9, [-1] * 9, 0.1, 1, np.log(100)]).astype("float64")

    def getPars(self, *args, **kwargs) -> np.ndarray:
        return np.hstack([self.kappa, self.gamma, self.a, self.sigma, self.tau]).astype("float64")

    def setPars(self, par) -> None:
        par = np.array(par, dtype="float64")
        self.kappa, self.gamma = par[0:9], par[9:18]
        self.a, self.sigma, self.tau = par[18], par[19], par[20]

    def transDiff(self, par=None):
        raise AttributeError("SeperableSpatialTemporalIDiffusion2D has no transDiff (half-angle classes only)")

    def print(self, par):
        return ("| κ = %2.2f" % (np.exp(par[0:9]).mean()) + ", γ = %2.2f" % (np.exp(par[9:18]).mean())
                + ", a = %2.2f" % (par[18]) + ", σ\t = %2.2f" % (np.exp(par[19])) + ", τ = %2.2f" % (np.exp(par[20])))   # (sic)


class _KronDQ:
    """``kron(Qt, dQs_i)`` / ``kron(dQt, Qs)`` of the reference's ``dQ`` list (``:84-121``) as an operator."""

    def __init__(self, model, st, spatial_op, dqt):
        self.model, self.st, self.spatial_op, self.dqt = model, st, spatial_op, dqt
        self._slots = None
        n = model.engine.n
        self.shape = (n, n)

    def slots(self):
        if self._slots is None:
            eng = self.model.engine
            if self.spatial_op is not None:
                self._slots = eng.fill_kron(self.spatial_op.slots(), *self.st["qt"])
            else:
                self._slots = eng.fill_kron(self.st["spatial"]["Q"], *self.dqt)
        return self._slots

    def __matmul__(self, X):
        eng = self.model.engine
        is_np = not isinstance(X, torch.Tensor)
        x = to_dev(np.asarray(X, dtype=np.float64) if is_np else X)
        one = x.dim() == 1
        y = eng.q_apply(self.slots(), x.reshape(eng.n, -1).contiguous())
        if one:
            y = y.reshape(-1)
        return to_host(y) if is_np else y

    def tocsc(self):
        return self.model.engine.to_scipy(self.slots())
