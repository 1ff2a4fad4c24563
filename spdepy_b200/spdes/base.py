"""SPDE model classes on the B200 engine -- one parametrised implementation behind the reference's
per-class API (``__init__(grid,[mod0],par,bc)``, ``getPars/setPars/initFit/setQ/makeQ/logLike/print/
Ah/Aw``; attributes ``Q, Q_fac, S, data, r, tau, type, mod0``).

What a class of the reference does with ``scipy.sparse`` + CHOLMOD per call
(``advection_diffusion2D.py:86-223``) happens here on the device:

  fields (host NumPy, as in the reference) -> K2 stencils -> K3 assembly into the slot layout
  -> K4/K5 supernodal Cholesky of Q and of Q + tau S^T S -> K6 logdet, K7 solves, K8 quadratic
  forms -> gradient: weights on the pattern of Q (Hutchinson probes via SDDMM, or the exact
  Takahashi inverse) -> K11 adjoint of the assembly -> one dot product per parameter.

``logLike`` keeps the reference signature and return convention (``-like/(nobs r)``, ``-g/(nobs r)``)
and adds two keyword arguments: ``probes=`` (inject the +-1 probe matrix instead of drawing it from
the global legacy RNG) and ``exact_grad=`` (Takahashi traces instead of the
Hutchinson estimator).
"""
from __future__ import annotations

import numpy as np
import torch
from scipy import sparse

from ..engine import COUNTERS, F64, Engine, Lazy, ScalarPool, nvtx_range, to_dev, to_host


class SPDE2D:
    # ---- configured by the concrete classes in spdes/__init__.py
    name = ""            # reference ``type`` prefix
    timed = False
    kvar = False         # spline kappa
    Hkind = "iso"        # "iso" | "aniso" | "ha"
    Hvar = False
    wkind = None         # None | "const" | "var"
    aflav = 1            # spde_combine_A flavour for A (0 spatial, 1 "sum", 2 "paren")
    divide = False       # Q / (dt*sigma) instead of (1/(dt*sigma)) * Q
    default_own = ()     # default own parameters (without mod0 block and tau)

    def __init__(self, grid, mod0=None, par=None, bc=3, ww=None) -> None:
        self.grid = grid
        self.ww = None if ww is None else np.ascontiguousarray(ww, dtype="float64")   # cov-advection: face velocities
        self.type = "%s-bc%d" % (self.name, bc)
        self.Q = None
        self.Q_fac = None
        self.data = None
        self.r = None
        self.S = None
        self.bc = bc
        self.mod0 = mod0
        self.Np = grid.Nbs2
        self.fitQ0 = True
        self._state = None
        self._obs = None
        M, N = grid.shape[0], grid.shape[1]
        self.engine = Engine.get(M, N, grid.T if self.timed else 1, bc)
        if self.timed and mod0 is None:
            raise ValueError("space-time models need an initial-field model mod0")
        if bc == 2 and not self.Hvar:
            raise ValueError("bc=2 with a constant diffusion tensor is undefined in the reference "
                             "(AcH_2D_b2.cpp:105 returns NaN); use a var-* model")
        if par is None or (self.wkind == "cov" and self.ww is None):
            # (the cov-advection classes of the reference never assemble in __init__, cov_advection_diffusion2D.py:26-28)
            self.setPars(self._default_par(joint=False) if par is None else par)
        else:
            self.setQ(par=par)

    # ------------------------------------------------------------------ parameters
    def _default_own(self):
        out = []
        for v in self.default_own:
            out.extend(v if isinstance(v, (list, tuple)) else [v])
        return out

    def _default_par(self, joint: bool):
        own = self._default_own()
        if joint and self.timed:
            return np.hstack([own, self.mod0.getPars()[:-1], np.log(100)]).astype("float64")
        return np.hstack([own, np.log(100)]).astype("float64")

    @property
    def n_own(self) -> int:
        nk = self.Np if self.kvar else 1
        nh = {"iso": 1, "aniso": 3, "ha": 3}[self.Hkind] * (self.Np if self.Hvar else 1)
        nw = {None: 0, "const": 2, "cov": 1, "var": 2 * self.Np}[self.wkind]
        return nk + nh + nw + (1 if self.timed else 0)

    def _split(self, par):
        nk = self.Np if self.kvar else 1
        nh = self.Np if self.Hvar else 1
        o = 0
        p = {"kappa": par[o:o + nk]}
        o += nk
        p["gamma"] = par[o:o + nh]
        o += nh
        if self.Hkind != "iso":
            p["vx"], p["vy"] = par[o:o + nh], par[o + nh:o + 2 * nh]
            o += 2 * nh
        if self.wkind == "const":
            p["w"] = par[o:o + 2]
            o += 2
        elif self.wkind == "var":
            p["w"] = par[o:o + 2 * self.Np]
            o += 2 * self.Np
        elif self.wkind == "cov":
            p["w"] = par[o:o + 1]
            o += 1
        if self.timed:
            p["sigma"] = par[o]
            o += 1
        return p

    def getPars(self, onlySelf=True) -> np.ndarray:
        if onlySelf or not self.timed:
            return np.hstack([self._own, self.tau]).astype("float64")
        return np.hstack([self._own, self.mod0.getPars()[:-1], self.tau]).astype("float64")

    def transDiff(self, par=None):
        """Half-angle classes only: the (gamma, vx, vy) of the equivalent ``gamma I + v v^T`` parametrisation, left in
        ``self.tgamma / tvx / tvy`` as the reference does.  Constant-coefficient form ``whittle_matern_ha2D.py:37-45``
        (it scales ``tvx, tvy`` with ``exp(par[0])``; reproduced), spline-field form ``advection_var_ha_diffusion2D.py:
        46-58``; ``VarWhittleMaternHa2D`` uses the constant form on its first four entries (``var_whittle_matern_ha2D.py:
        37-45``; reproduced)."""
        if self.Hkind != "ha":
            raise AttributeError("%s has no transDiff (half-angle classes only)" % type(self).__name__)
        par = self.getPars() if par is None else np.asarray(par, dtype="float64")
        if self.Hvar and self.timed:
            g0, vx, vy = np.exp(self.grid.evalB(par[9:18])), self.grid.evalB(par[18:27]), self.grid.evalB(par[27:36])
            g1 = g0
        else:
            # the scale under the square roots is exp(par[0]) in four classes and exp(par[1]) in the cov- / var-advection
            # ones (cov_advection_ha_diffusion2D.py:54-56, var_advection_ha_diffusion2D.py:53-55)
            g0, g1, vx, vy = np.exp(par[1]), np.exp(par[1] if self.wkind in ("cov", "var") else par[0]), par[2], par[3]
        aV = np.sqrt(vx ** 2 + vy ** 2)
        cosh_aV = (np.exp(aV) + np.exp(-aV)) / 2
        sinh_aV = (np.exp(aV) - np.exp(-aV)) / 2
        self.tgamma = g0 * (cosh_aV - sinh_aV)
        self.tvx = np.sqrt(g1 * sinh_aV / aV * (vx + aV))
        self.tvy = np.sqrt(g1 * sinh_aV / aV * (-vx + aV))

    def setPars(self, par) -> None:
        par = np.array(par, dtype="float64")
        self._own = par[:self.n_own].copy()
        p = self._split(par)
        sc = (lambda v: v if v.size > 1 else float(v[0]))
        self.kappa, self.gamma = sc(p["kappa"]), sc(p["gamma"])
        if "vx" in p:
            self.vx, self.vy = sc(p["vx"]), sc(p["vy"])
        if self.wkind == "const":
            self.wx, self.wy = float(p["w"][0]), float(p["w"][1])
        elif self.wkind == "var":
            self.wx, self.wy = p["w"][:self.Np], p["w"][self.Np:]
        elif self.wkind == "cov":
            self.lamb = float(p["w"][0])
        if self.timed:
            self.sigma = float(p["sigma"])
            if par.size > self.n_own + 1:
                # as the reference: the initial-field model is re-assembled with its block of the joint vector
                # (mod0.setQ(par=par[7:]), advection_diffusion2D.py:46-47), so that later own-parameter calls
                # (setQ(), logLike(..., fitQ0=False), Model.setModel) read the new Q0 and not a stale one
                self.mod0.setQ(par=par[self.n_own:])
        self.tau = par[-1]
        if not self.timed:
            self.sigma = np.log(np.sqrt(1 / np.exp(self.tau)))

    def initFit(self, data, **kwargs):
        data = np.asarray(data, dtype="float64")
        assert data.shape[0] <= self.grid.n
        if kwargs.get("fitQ0") is not None:
            assert type(kwargs.get("fitQ0")) is bool
            self.fitQ0 = kwargs.get("fitQ0")
        par = self._default_par(joint=self.fitQ0)
        self.data = data
        self.r = data.shape[1] if data.ndim == 2 else 1
        if kwargs.get("ww") is not None:
            self.ww = np.ascontiguousarray(kwargs.get("ww"), dtype="float64")
        idx = kwargs.get("idx")
        self.S = self.grid.getS(idxs=idx)
        self._set_obs(self.grid.obs_nodes(idx))
        return par

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
        self.Q, self.Q_fac, _ = self.makeQ(par=np.asarray(par, dtype="float64"), grad=False)
        self.S = self.grid.getS()
        self._set_obs(self.grid.obs_nodes())

    def print(self, par):
        p = self._split(np.asarray(par, dtype="float64"))
        s = "| κ = %2.2f" % np.exp(p["kappa"]).mean() + ", γ = %2.2f" % np.exp(p["gamma"]).mean()
        if "vx" in p:
            s += ", vx = %2.2f" % np.mean(p["vx"]) + ", vy = %2.2f" % np.mean(p["vy"])
        if self.wkind == "cov":
            # the constant-coefficient classes print the multiplier itself, the spline-field ones its exponential
            # (cov_advection_diffusion2D.py:85, cov_advection_var_diffusion2D.py:85)
            s += ", Λ = %2.2f" % (np.exp(p["w"][0]) if self.Hvar else p["w"][0])
        elif self.wkind is not None:
            h = p["w"].size // 2
            s += ", wx = %2.2f" % np.mean(p["w"][:h]) + ", wy = %2.2f" % np.mean(p["w"][h:])
        if self.timed:
            s += ", σ = %2.2f" % np.exp(p["sigma"])
        s += (", τ = %2.2f" if self.timed else ",τ = %2.2f") % np.exp(par[-1])       # (sic, whittle_matern2D.py print)
        if self.timed and par.size > self.n_own + 1:
            s += "\n Q0: " + self.mod0.print(par[self.n_own:])[:-10]
        return s

    # ------------------------------------------------------------------ fields (host NumPy, like the reference)
    def _H_and_dirs(self, p, want_dirs):
        """Diffusion tensor and, per diffusion parameter in the reference's order, dH/dtheta_i
        (``advection_diffusion2D.py:100,131-155``; half-angle ``whittle_matern_ha2D.py:73-106``;
        spline fields ``var_advection_var_diffusion2D.py:91-99,131-160``)."""
        g = self.grid
        dirs = []
        I2 = np.eye(2)
        if not self.Hvar:
            gam = np.exp(p["gamma"][0])
            if self.Hkind == "iso":
                H = gam * I2
                dirs = [gam * I2]
            elif self.Hkind == "aniso":
                vv = np.array([p["vx"][0], p["vy"][0]])
                H = gam * I2 + np.outer(vv, vv)
                dirs = [gam * I2]
                for e in (np.array([1.0, 0.0]), np.array([0.0, 1.0])):
                    dirs.append(np.outer(e, vv) + np.outer(vv, e))
            else:
                vx, vy = p["vx"][0], p["vy"][0]
                aV = np.sqrt(vx ** 2 + vy ** 2)
                mV = np.array([[vx, vy], [vy, -vx]])
                ch, sh = (np.exp(aV) + np.exp(-aV)) / 2, (np.exp(aV) - np.exp(-aV)) / 2
                H = gam * (ch * I2 + sh / aV * mV)
                dirs = [H,
                        gam / aV * (vx * sh * I2 + vx / aV * (ch - sh / aV) * mV + sh * np.array([[1, 0], [0, -1]])),
                        gam / aV * (vy * sh * I2 + vy / aV * (ch - sh / aV) * mV + sh * np.array([[0, 1], [1, 0]]))]
            return H, dirs
        gam = np.exp(g.evalBH(par=p["gamma"]))                      # (Ns, 4)
        H = I2 * (np.stack([gam, gam], axis=2))[:, :, :, None]
        self._face_fields = {"gam": gam}
        if self.Hkind == "aniso":
            vv = np.stack([g.evalBH(p["vx"]), g.evalBH(p["vy"])], axis=2)
            H = H + vv[:, :, :, None] * vv[:, :, None, :]
            self._face_fields.update(vx=vv[:, :, 0], vy=vv[:, :, 1])
        elif self.Hkind == "ha":        # advection_var_ha_diffusion2D.py:104-113
            vx, vy = g.evalBH(p["vx"]), g.evalBH(p["vy"])
            aV = np.sqrt(vx ** 2 + vy ** 2)
            mV = np.array([[vx, vy], [vy, -vx]]).T.swapaxes(0, 1)
            ch, sh = (np.exp(aV) + np.exp(-aV)) / 2, (np.exp(aV) - np.exp(-aV)) / 2
            H = (gam * ch)[:, :, None, None] * I2 + (gam * sh / aV)[:, :, None, None] * mV
            self._face_fields.update(vx=vx, vy=vy)
            if want_dirs:
                for i in range(self.Np):
                    dg = g.bsH[:, :, i] * gam
                    dirs.append((dg * ch)[:, :, None, None] * I2 + (dg * sh / aV)[:, :, None, None] * mV)
                for comp in (0, 1):
                    for i in range(self.Np):
                        dv, z = g.bsH[:, :, i], vx * 0
                        dmV = (np.array([[dv, z], [z, -dv]]) if comp == 0 else np.array([[z, dv], [dv, -z]])).T.swapaxes(0, 1)
                        vc = vx if comp == 0 else vy
                        dirs.append((gam * dv * vc / aV)[:, :, None, None] * (sh[:, :, None, None] * I2 + ((ch - sh / aV) / aV)[:, :, None, None] * mV)
                                    + (gam * sh / aV)[:, :, None, None] * dmV)
            return H, dirs
        if want_dirs:
            for i in range(self.Np):
                dg = g.bsH[:, :, i] * gam
                dirs.append(I2 * (np.stack([dg, dg], axis=2)[:, :, :, None]))
            if self.Hkind == "aniso":
                zero = np.zeros_like(gam)
                for comp in (0, 1):
                    for i in range(self.Np):
                        b = g.bsH[:, :, i]
                        dv = np.stack([b, zero], axis=2) if comp == 0 else np.stack([zero, b], axis=2)
                        dirs.append(vv[:, :, :, None] * dv[:, :, None, :] + dv[:, :, :, None] * vv[:, :, None, :])
        return H, dirs

    # ------------------------------------------------------------------ assembly on the device
    def _assemble(self, par, need_Q0_state=False):
        """theta -> device state: kappa, A9, Q (slot layout) (+ mod0 state)."""
        g, eng = self.grid, self.engine
        par = np.asarray(par, dtype="float64")
        p = self._split(par)
        V = g.V
        dt = g.dt if self.timed else 0.0
        kap = np.exp(g.evalB(par=p["kappa"])) if self.kvar else np.exp(p["kappa"][:1])
        H, _ = self._H_and_dirs(p, want_dirs=False)
        st = {"par": par, "p": p, "V": V, "dt": dt, "kappa": to_dev(kap), "H": H}
        if self.Hvar:
            st["faces"] = {k: to_dev(np.ascontiguousarray(v)) for k, v in self._face_fields.items()}
        ah = eng.ah_stencil(g.hx, g.hy, to_dev(H), self.Hvar)
        aw = None
        if self.wkind == "const":
            st["ws"] = np.array(p["w"], dtype="float64")
            aw = eng.aw_stencil(g.hx, g.hy, to_dev(st["ws"]), None, False, 3, False)
        elif self.wkind == "var":
            st["ws"] = np.ascontiguousarray(g.evalAdv(p["w"]))
            aw = eng.aw_stencil(g.hx, g.hy, to_dev(st["ws"]), None, True, 3, True)
        elif self.wkind == "cov":        # ws = lambda * ww  (cov_advection_diffusion2D.py:101)
            if self.ww is None:
                raise ValueError("cov-advection models need the face velocities ww (initFit(..., ww=...))")
            st["ws"] = np.ascontiguousarray(p["w"][0] * self.ww)
            aw = eng.aw_stencil(g.hx, g.hy, to_dev(st["ws"]), None, True, 3, False)
        st["A9"] = eng.combine_A(self.aflav if self.timed else 0, V, dt, st["kappa"], ah, aw)
        if not self.timed:
            st["Q"] = eng.atda(st["A9"], st["kappa"], V, 0)
            st["sigma"] = 1.0
            return st
        sigma = float(np.exp(p["sigma"]))
        st["sigma"] = sigma
        joint = par.size > self.n_own + 1
        st["joint"] = joint
        if joint:
            st0 = self.mod0._assemble(par[self.n_own:])
        else:
            if self.mod0._state is None:
                self.mod0.setQ()
            st0 = self.mod0._state
        st["mod0"] = st0
        AtDA = eng.atda(st["A9"], st["kappa"], V, 1)
        st["AtDA"] = AtDA
        st["Q"] = eng.fill_spacetime(AtDA, st["A9"], st["kappa"], V, st0["Q"], sigma, dt, self.divide)
        return st

    def makeQ(self, par, grad=True):
        """Returns ``(Q, Q_fac, dQ)`` like the reference.  ``Q`` is a SciPy CSC matrix exported from
        the device, ``Q_fac`` a :class:`Factor` of it.  The reference materialises one n x n sparse
        ``dQ_i`` per parameter (``advection_diffusion2D.py:119-182``); here the derivative enters
        ``logLike`` through the adjoint of the assembly and is never formed, so ``dQ`` is a list of
        callables ``dQ[i](X)`` evaluated by directional differencing of the assembly."""
        st = self._assemble(np.asarray(par, dtype="float64"))
        self._state = st
        fac = self.engine.factorize(0, st["Q"])
        Q = self.engine.to_scipy(st["Q"])
        return Q, fac, (self._lazy_dQ(st) if grad else None)

    # ------------------------------------------------------------------ dQ/dtheta_i as lazy operators
    def _Q_from(self, st, A9, kappa):
        """Precision (slot layout) for a given operator stencil and kappa field, without the initial-field block."""
        eng = self.engine
        if not self.timed:
            return eng.atda(A9, kappa, st["V"], 0)
        AtDA = eng.atda(A9, kappa, st["V"], 1)
        zero0 = torch.zeros(25 * eng.Ns, dtype=F64, device=A9.device)
        return eng.fill_spacetime(AtDA, A9, kappa, st["V"], zero0, st["sigma"], st["dt"], self.divide)

    def _direction_stencils(self, st):
        """(kind, payload) per own parameter, in the reference's dQ order."""
        g, eng = self.grid, self.engine
        V, dt = st["V"], st["dt"]
        out = []
        nk = self.Np if self.kvar else 1
        for i in range(nk):
            out.append(("kappa", i))
        _, dirs = self._H_and_dirs(st["p"], want_dirs=True)
        for dH in dirs:
            ah = eng.ah_stencil(g.hx, g.hy, to_dev(dH), self.Hvar)
            out.append(("dA", eng.combine_A(3 if self.timed else 5, V, dt, st["kappa"], ah, None)))
        if self.wkind is not None:
            wsd = to_dev(st["ws"])
            if self.wkind == "const":
                for d in (1, 2):
                    aw = eng.aw_stencil(g.hx, g.hy, wsd, None, False, d, False)
                    out.append(("dA", eng.combine_A(4, V, dt, st["kappa"], None, aw)))
            elif self.wkind == "cov":
                aw = eng.aw_stencil(g.hx, g.hy, to_dev(self.ww), None, True, 3, False)
                out.append(("dA", eng.combine_A(4, V, dt, st["kappa"], None, aw)))
            else:
                for i in range(2 * self.Np):
                    dpar = np.zeros(2 * self.Np)
                    dpar[i] = 1
                    dws = to_dev(np.ascontiguousarray(g.evalAdv(dpar)))
                    aw = eng.aw_stencil(g.hx, g.hy, wsd, dws, True, 1 if i < self.Np else 2, True)
                    out.append(("dA", eng.combine_A(4, V, dt, st["kappa"], None, aw)))
        if self.timed:
            out.append(("sigma", None))
        return out

    def _dQ_slots(self, st, kind, payload):
        """Slot array of one dQ/dtheta_i.  Q is exactly quadratic in the operator stencil A and linear in
        Qs = kappa^2 V, so symmetric / forward differences of the *assembly kernels* with O(1) steps
        reproduce the reference's analytic dQ (advection_diffusion2D.py:119-182) to rounding."""
        eng = self.engine
        A9, kap = st["A9"], st["kappa"]
        Ns = eng.Ns
        if kind == "sigma":
            Qnp = self._Q_from(st, A9, kap)
            return -Qnp
        if kind == "dA":
            dA = payload
            s = float(A9.abs().max() / dA.abs().max().clamp_min(1e-300))       # comparable magnitudes: well conditioned
            return (self._Q_from(st, A9 + s * dA, kap) - self._Q_from(st, A9 - s * dA, kap)) / (2 * s)
        # log kappa (spline coefficient i or the scalar): dA on the centre slot + the linear Qs dependence
        bs = self._bs_dev()[:, payload] if self.kvar else torch.ones(1, dtype=F64, device=A9.device)
        dA = torch.zeros_like(A9)
        dA[4 * Ns:5 * Ns] = st["V"] * kap * bs * (st["dt"] if self.timed else 1.0)
        s = float(A9.abs().max() / dA.abs().max().clamp_min(1e-300))
        out = (self._Q_from(st, A9 + s * dA, kap) - self._Q_from(st, A9 - s * dA, kap)) / (2 * s)
        if self.timed:
            h = 0.25
            kap2 = kap * torch.sqrt(1 + 2 * h * bs)          # Qs -> Qs (1 + 2 h bs) = Qs + h dQs
            out = out + (self._Q_from(st, A9, kap2) - self._Q_from(st, A9, kap)) / h
        return out

    def _lazy_dQ(self, st):
        ops = [LazyDQ(self, st, kind, payload) for kind, payload in self._direction_stencils(st)]
        if self.timed and st["joint"]:
            for op0 in self.mod0._lazy_dQ(st["mod0"]):
                ops.append(LazyDQ(self, st, "q0", op0))
        return ops

    # public stencil wrappers with the reference's return type (advection_diffusion2D.py:226-260)
    def _stencil_to_csc(self, a9):
        eng = self.engine
        a9 = a9.cpu().numpy().reshape(9, eng.Ns)
        k = np.arange(eng.Ns)
        i, j = k % eng.M, k // eng.M
        rows, cols, vals = [], [], []
        for s in range(9):
            ii, jj = i + (s % 3 - 1), j + (s // 3 - 1)
            if self.bc == 2:
                ii, jj = ii % eng.M, jj % eng.N
                ok = np.ones(eng.Ns, bool)
            else:
                ok = (ii >= 0) & (ii < eng.M) & (jj >= 0) & (jj < eng.N)
            rows.append(k[ok]); cols.append((jj * eng.M + ii)[ok]); vals.append(a9[s][ok])
        return sparse.csc_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(eng.Ns, eng.Ns))

    def Ah(self, Hs) -> sparse.csc_matrix:
        Hs = np.array(Hs, dtype="float64")
        return self._stencil_to_csc(self.engine.ah_stencil(self.grid.hx, self.grid.hy, to_dev(Hs), Hs.ndim == 4))

    def Aw(self, ws, dws=None, diff=3) -> sparse.csc_matrix:
        ws = np.array(ws, dtype="float64")
        face = ws.ndim == 2
        d = None if dws is None else to_dev(np.array(dws, dtype="float64"))
        return self._stencil_to_csc(self.engine.aw_stencil(self.grid.hx, self.grid.hy, to_dev(ws), d, face, diff, face))

    def setClib(self) -> None:      # the reference compiles / loads its C++ here; nothing to do
        return None

    # ------------------------------------------------------------------ scalar results: eager, or collected on the device
    _pool = None         # ScalarPool of the evaluation in flight (logLike), None = every reduction returns a host float

    def _dot(self, X, Y):
        return Engine.dot(X, Y) if self._pool is None else self._pool.dot(X, Y)

    def _wdot(self, X, Y, w):
        return Engine.wdot(X, Y, w) if self._pool is None else self._pool.wdot(X, Y, w)

    def _resid(self, data, mu, obs):
        return Engine.residual_ss(data, mu, obs) if self._pool is None else self._pool.residual_ss(data, mu, obs)

    def _logdet(self, eng, which):
        return eng.logdet(which) if self._pool is None else self._pool.logdet(eng, which)

    def _gemv_t(self, B, u, scale=1.0):
        if self._pool is None:
            return to_host(scale * Engine.gemv_t(B, u)).tolist()
        return [scale * v for v in self._pool.gemv_t(B, u)]

    # ------------------------------------------------------------------ gradient contraction (K11)
    def _grad_from_weights(self, st, W, prior=None):
        """sum(W .* dQ_i) for every own parameter i (and the mod0 block when fitted jointly),
        from the adjoint of the assembly; parameter order as ``advection_diffusion2D.py:119-182``.
        ``prior`` (space-time models): the time-collapsed prior term ``c * d logdet Q / d theta`` given as
        ``{"c": c, "ZB": selected inverse of B, "Z0": selected inverse of Q0}``, see :meth:`_prior_collapsed`."""
        g, eng = self.grid, self.engine
        Ns = eng.Ns
        V, dt, sigma = st["V"], st["dt"], st["sigma"]
        GA, Gq, GQ0 = eng.assembly_adjoint(W, st["A9"], st["kappa"], V, sigma, dt, self.timed)
        s_q0 = self._dot(GQ0, st["mod0"]["Q"]) if self.timed else 0.0
        s_prior_sigma = 0.0
        if prior is not None:
            # c*logdet Q = c*logdet Q0 + c*(T-1) (logdet B - Ns log(dt sigma)): weights c (T-1) Z_B on B, c Z_0 on Q0
            c = prior["c"]
            GA2, Gq2 = self.engine2d.assembly_adjoint_B(prior["ZB"] * (c * (eng.T - 1)), st["A9"], st["kappa"], V)
            GA = GA + GA2
            Gq = Gq + Gq2
            if prior.get("Z0") is not None:
                GQ0 = GQ0 + c * prior["Z0"]
            s_prior_sigma = -c * (eng.T - 1) * Ns
        out = []
        # kappa
        kap = st["kappa"]
        GAc = GA[4 * Ns:5 * Ns]
        if self.timed:
            qs = ((V * kap) * (1.0 / V)) * (V * kap)
            u = (V * dt) * kap * GAc + 2.0 * qs * Gq
        else:
            u = V * kap * GAc
        if self.kvar:
            out.extend(self._gemv_t(self._bs_dev(), u.contiguous()))
        else:
            out.append(self._dot(u.contiguous(), torch.ones_like(u)))
        # diffusion
        sgn = -dt if self.timed else -1.0
        wsd = to_dev(st["ws"]) if self.wkind is not None else None
        if self.Hvar:
            # spline fields: adjoint of the face stencil, then one GEMV with the face basis per parameter block
            GH, GdG = eng.stencil_adjoint(g.hx, g.hy, GA, True, wsd if self.wkind == "var" else None)
            GH = GH.view(8, Ns)
            f = st["faces"]
            D = torch.stack([GH[0], GH[1], GH[6], GH[7]], dim=1)        # dS/d(H00 at W,E), d(H11 at S,N)
            bsH = self._bsH_dev()
            out.extend(self._gemv_t(bsH, (f["gam"] * D).reshape(-1).contiguous(), sgn))
            if self.Hkind == "ha":
                # half-angle parametrisation H = gamma (cosh|v| I + sinh|v|/|v| [[vx,vy],[vy,-vx]])
                out = out[:-self.Np]                                    # gamma block recomputed below (dH/dlog gamma = H)
                O = torch.stack([GH[2], GH[3], GH[4], GH[5]], dim=1)
                vx, vy, gam = f["vx"], f["vy"], f["gam"]
                a = torch.sqrt(vx * vx + vy * vy)
                ch, sh = torch.cosh(a), torch.sinh(a)
                s = sh / a
                ds = (ch - s) / a
                sg = torch.tensor([1.0, 1.0, -1.0, -1.0], dtype=F64, device=a.device)   # H00 on W,E ; H11 on S,N
                u_g = gam * ((ch + sg * s * vx) * D + s * vy * O)
                u_vx = gam * ((sh * vx / a + sg * (ds * vx / a * vx + s)) * D + ds * vx / a * vy * O)
                u_vy = gam * ((sh * vy / a + sg * (ds * vy / a * vx)) * D + (ds * vy / a * vy + s) * O)
                for u in (u_g, u_vx, u_vy):
                    out.extend(self._gemv_t(bsH, u.reshape(-1).contiguous(), sgn))
            elif self.Hkind == "aniso":
                O = torch.stack([GH[2], GH[3], GH[4], GH[5]], dim=1)    # dS/d(H10 at W,E), d(H01 at S,N)
                vx, vy = f["vx"], f["vy"]
                u_vx = torch.cat([2 * vx[:, :2] * D[:, :2] + vy[:, :2] * O[:, :2], vy[:, 2:] * O[:, 2:]], dim=1)
                u_vy = torch.cat([vx[:, :2] * O[:, :2], vx[:, 2:] * O[:, 2:] + 2 * vy[:, 2:] * D[:, 2:]], dim=1)
                out.extend(self._gemv_t(bsH, u_vx.reshape(-1).contiguous(), sgn))
                out.extend(self._gemv_t(bsH, u_vy.reshape(-1).contiguous(), sgn))
        else:
            GdG = None
            _, dirs = self._H_and_dirs(st["p"], want_dirs=True)
            for dH in dirs:
                ah = eng.ah_stencil(g.hx, g.hy, to_dev(dH), self.Hvar)
                out.append(sgn * self._dot(GA, ah))
        # advection
        if self.wkind == "const":
            for d in (1, 2):
                out.append(dt * self._dot(GA, eng.aw_stencil(g.hx, g.hy, wsd, None, False, d, False)))
        elif self.wkind == "cov":        # dA = Aw(ww) * dt   (cov_advection_diffusion2D.py:60,158-159)
            out.append(dt * self._dot(GA, eng.aw_stencil(g.hx, g.hy, to_dev(self.ww), None, True, 3, False)))
        elif self.wkind == "var":
            if GdG is None:
                _, GdG = eng.stencil_adjoint(g.hx, g.hy, GA, False, wsd)
            BAx, BAy = self._bsA_dev()
            out.extend(self._gemv_t(BAx, GdG[:, [0, 2]].reshape(-1).contiguous(), dt))
            out.extend(self._gemv_t(BAy, GdG[:, [1, 3]].reshape(-1).contiguous(), dt))
        if self.timed:
            # log sigma: dQ = -(Q - blockdiag(Q0-part))   (advection_diffusion2D.py:170-175)
            s_total = self._dot(W, st["Q"])
            out.append(-(s_total - s_q0) + s_prior_sigma)
            if st["joint"]:
                self.mod0._pool = self._pool
                try:
                    out.extend(self.mod0._grad_from_weights(st["mod0"], GQ0))
                finally:
                    self.mod0._pool = None
        return out

    def _bsH_dev(self):
        if getattr(self, "_bsH_cache", None) is None:
            self._bsH_cache = to_dev(np.ascontiguousarray(self.grid.bsH.reshape(-1, self.Np)))
        return self._bsH_cache

    def _bsA_dev(self):
        """Effective advection bases: d evalAdv / d theta_i = advBound(bsA[:, f, i]) (spat2Dtemp_regular_mesh.py:267-275)."""
        if getattr(self, "_bsA_cache", None) is None:
            g = self.grid
            B = g.advBound(g.bsA.reshape(g.bsA.shape[0], -1)).reshape(-1, 4, self.Np)
            self._bsA_cache = (to_dev(np.ascontiguousarray(B[:, [0, 2], :].reshape(-1, self.Np))),
                               to_dev(np.ascontiguousarray(B[:, [1, 3], :].reshape(-1, self.Np))))
        return self._bsA_cache

    def _bs_dev(self):
        if getattr(self, "_bs_cache", None) is None:
            self._bs_cache = to_dev(np.ascontiguousarray(self.grid.bs))
        return self._bs_cache

    # ------------------------------------------------------------------ time-collapsed prior
    collapse_prior = True    # False: factorise the 3-D prior precision as the reference does (validation)

    @property
    def engine2d(self) -> Engine:
        """Engine of one time slice (the 5x5 pattern of B = A^T (Qs/V^2) A)."""
        return Engine.get(self.engine.M, self.engine.N, 1, self.bc)

    def _prior_collapsed(self, st, want_grad):
        """logdet of the space-time prior and the selected inverses that carry its gradient, from two 2-D
        factorisations.  The prior of ``advection_diffusion2D.py:112-116`` is ``Q = blockdiag(Q0, 0, ...) +
        c G^T G`` with ``c = 1/(dt sigma)`` and ``G`` block bidiagonal (rows ``[-Qs^1/2, Qs^1/2 V^-1 A]``), so

            logdet Q = logdet Q0 + (T-1) (Ns log c + logdet B),    B = A^T (Qs/V^2) A  (``spde_atda`` mode 1),

        and ``tr(Q^-1 dQ_i) = tr(Q0^-1 dQ0_i) + (T-1) (tr(B^-1 dB_i) - Ns dlog(sigma)_i)``."""
        e2, e0 = self.engine2d, self.mod0.engine
        T, Ns = self.engine.T, self.engine.Ns
        # store 1 of the slice engine holds B, store 0 of mod0's engine holds Q0 (the two may be one engine)
        e0.factorize_async(0, st["mod0"]["Q"])
        e2.factorize_async(1, st["AtDA"])
        e0.factor_wait(0)
        e2.factor_wait(1)
        ld0, ldB = self._logdet(e0, 0), self._logdet(e2, 1)
        out = {"logdet": ld0 + (T - 1) * (ldB - Ns * np.log(st["dt"] * st["sigma"])), "logdetQ0": ld0, "logdetB": ldB}
        if want_grad:
            out["ZB"] = e2.selinv(1)
            out["Z0"] = e0.selinv(0) if st["joint"] else None
        return out

    # ------------------------------------------------------------------ likelihood (advection_diffusion2D.py:187-223)
    def logLike(self, par, nh1=100, grad=True, probes=None, exact_grad=False):
        """Every scalar of the evaluation (log-determinants, quadratic forms, traces, the per-parameter contractions)
        is reduced into one device vector (:class:`ScalarPool`) and read back ONCE, when the likelihood and the gradient
        are put together at the end -- the arithmetic below is written on lazy scalars exactly as the reference writes it
        on NumPy floats (``advection_diffusion2D.py:198-223``)."""
        self._pool = ScalarPool()
        try:
            return self._logLike(par, nh1, grad, probes, exact_grad)
        finally:
            self._pool = None

    def _finish(self, like, norm, gi=None, g_last=None, npar=0):
        """The one device-to-host read of the evaluation (lazy scalars -> floats) and the reference's return convention,
        ``-like / (nobs r)`` and ``-grad / (nobs r)``."""
        for k, v in list(self.last.items()):
            if isinstance(v, Lazy):
                self.last[k] = float(v)
        like = float(like)
        if gi is None:
            return -like / norm
        g_par = np.zeros(npar)
        g_par[:len(gi)] = [float(v) for v in gi]
        g_par[-1] = float(g_last)
        return -like / norm, -g_par / norm

    def _logLike(self, par, nh1, grad, probes, exact_grad):
        eng = self.engine
        par = np.asarray(par, dtype="float64")
        if self._obs is None or self.data is None:
            raise RuntimeError("call initFit(data, idx=...) first")
        r, nobs = self.r, self._obs["nobs"]
        obs, cnt = self._obs["nodes"], self._obs["cnt"]
        data = self._obs.get("data")            # device-resident copy (bench.py's kernel-only arm)
        if data is None:
            data = to_dev(self.data.reshape(nobs, r))
        tau = float(np.exp(par[-1]))
        with nvtx_range("spde.assemble (K2/K3)"):
            st = self._assemble(par)
        self._state = st
        Q = st["Q"]
        # space-time prior: its determinant (and hence every prior trace of the gradient) factorises over
        # time into 2-D problems, so only the posterior precision needs the 3-D factorisation.  The
        # Hutchinson estimator solves with Q itself and keeps the 3-D factor of the prior.
        collapsed = self.timed and self.collapse_prior and (exact_grad or not grad)
        if eng.use_streamed(stores=1 if collapsed else 2):
            if not collapsed:
                raise NotImplementedError("meshes beyond device memory (streamed evaluation) need the time-collapsed prior: "
                                          "a space-time model with exact_grad=True or grad=False")
            return self._logLike_streamed(par, st, data, obs, cnt, tau, grad)
        if collapsed:
            with nvtx_range("spde.factorize Q_c (K4/K5) + collapsed prior"):
                eng.factorize_async(1, Q, cnt, tau)
                prior = self._prior_collapsed(st, want_grad=grad)
                eng.factor_wait(1)
            ldQ = prior["logdet"]
        else:
            prior = None
            eng.factorize_async(0, Q)               # the two factorisations run concurrently
            eng.factorize_async(1, Q, cnt, tau)
            if grad and not exact_grad and probes is None:
                # the reference's draw (advection_diffusion2D.py:200), from the same global legacy stream, made while the
                # device factorises; the +-1 matrix crosses PCIe as one byte per entry and is widened on the device
                r8 = np.random.randint(1, 3, self.grid.n * nh1).astype(np.int8)
                COUNTERS["h2d"] += r8.nbytes
                probes = (2.0 * torch.from_numpy(r8).to(device=Q.device, non_blocking=False).to(F64) - 3.0).reshape(self.grid.n, nh1)
            eng.factor_wait(0)
            eng.factor_wait(1)
            ldQ = self._logdet(eng, 0)
        ldQc = self._logdet(eng, 1)
        overlap = grad and exact_grad and collapsed
        if overlap:
            with nvtx_range("spde.selinv start (K10)"):
                eng.selinv_start(1)     # the Takahashi pass runs beside the (latency-bound) solve and reductions below
        with nvtx_range("spde.solve mu_c + reductions (K7/K8)"):
            mu_c = eng.solve(1, eng.scatter_obs(data, obs, tau))          # Q_c^-1 S^T data tau
            quad = self._dot(mu_c, eng.q_apply(Q, mu_c))
            resid = self._resid(data, mu_c, obs)
        like = 1 / 2 * ldQ * r + nobs * r * np.log(tau) / 2 - 1 / 2 * ldQc * r - 1 / 2 * quad - tau / 2 * resid
        self.last = {"mu_c": mu_c, "logdetQ": ldQ, "logdetQc": ldQc, "quad": quad, "resid": resid}
        if not grad:
            return self._finish(like, nobs * r)
        if exact_grad and collapsed:
            W = eng.selinv_fetch(1)
            nd = eng.nslots // 2
            tr_tau = self._dot(cnt, W[nd * eng.n:(nd + 1) * eng.n]) * tau
            W *= -0.5 * r
            prior["c"] = 0.5 * r
        elif exact_grad:
            Z, Zc = eng.selinv_pair()
            nd = eng.nslots // 2
            tr_tau = self._dot(cnt, Zc[nd * eng.n:(nd + 1) * eng.n].contiguous()) * tau
            W = (Z - Zc) * (0.5 * r)
            del Z, Zc
        else:
            if probes is None:
                probes = (2 * np.random.randint(1, 3, self.grid.n * nh1) - 3).reshape(self.grid.n, nh1)
            Vp = probes if isinstance(probes, torch.Tensor) else to_dev(np.asarray(probes, dtype=np.float64))
            nh1 = Vp.shape[1]
            TrQ = eng.solve(0, Vp.clone())
            TrQc = eng.solve(1, Vp.clone())
            a = 0.5 * r / nh1
            W = eng.sddmm(TrQ, Vp, a)
            W = eng.sddmm(TrQc, Vp, -a, W)
            tr_tau = self._wdot(TrQc, Vp, cnt) * tau / nh1
        with nvtx_range("spde.gradient contraction (K9/K11)"):
            W = eng.sddmm(mu_c, mu_c, -0.5, W)
            gi = self._grad_from_weights(st, W, prior)
            g_last = nobs * r / 2 - 1 / 2 * tr_tau * r - tau / 2 * resid
            return self._finish(like, nobs * r, gi, g_last, par.size)


    def _logLike_streamed(self, par, st, data, obs, cnt, tau, grad):
        """``logLike`` for meshes whose posterior factor does not fit in HBM (``Engine.streamed_eval``): the prior
        terms come from the 2-D factorisations of the time-collapsed prior, the posterior precision is factorised
        depth-first.  Without the gradient only the forward pass runs and the quadratic terms use
        ``mu^T Q mu + tau |y - S mu|^2 = tau y^T y - b^T Q_c^-1 b = tau y^T y - |L^-1 P b|^2`` (b = tau S^T y);
        with it, the backward pass returns the conditional mean and the selected inverse as the in-core path."""
        eng = self.engine
        r, nobs = self.r, self._obs["nobs"]
        Q = st["Q"]
        prior = self._prior_collapsed(st, want_grad=grad)
        ldQ = prior["logdet"]
        b = eng.scatter_obs(data, obs, tau)
        if not grad:
            ldQc, y, _ = eng.streamed_eval(Q, cnt, tau, X=b, mode=1 | 4, selinv=False)
            yy = self._dot(data, data)
            half = 0.5 * (tau * yy - self._dot(y, y))
            like = 1 / 2 * ldQ * r + nobs * r * np.log(tau) / 2 - 1 / 2 * ldQc * r - half
            self.last = {"mu_c": None, "logdetQ": ldQ, "logdetQc": ldQc}
            return self._finish(like, nobs * r)
        ldQc, mu_c, W = eng.streamed_eval(Q, cnt, tau, X=b, mode=15, selinv=True)
        quad = self._dot(mu_c, eng.q_apply(Q, mu_c))
        resid = self._resid(data, mu_c, obs)
        like = 1 / 2 * ldQ * r + nobs * r * np.log(tau) / 2 - 1 / 2 * ldQc * r - 1 / 2 * quad - tau / 2 * resid
        self.last = {"mu_c": mu_c, "logdetQ": ldQ, "logdetQc": ldQc, "quad": quad, "resid": resid}
        nd = eng.nslots // 2
        tr_tau = self._dot(cnt, W[nd * eng.n:(nd + 1) * eng.n]) * tau
        if getattr(self, "check_selinv", False):
            # size-independent checksum of the selected inverse: the pattern of Q_c is inside the extracted pattern,
            # so sum_e (Q_c)_e Z_e = tr(Q_c Q_c^-1) = n exactly
            self.last["selinv_trace_over_n"] = (self._dot(Q, W) + tr_tau) / eng.n
        W *= -0.5 * r
        prior["c"] = 0.5 * r
        W = eng.sddmm(mu_c, mu_c, -0.5, W)
        gi = self._grad_from_weights(st, W, prior)
        g_last = nobs * r / 2 - 1 / 2 * tr_tau * r - tau / 2 * resid
        return self._finish(like, nobs * r, gi, g_last, par.size)


class LazyDQ:
    """One ``dQ/dtheta_i`` of the reference's ``dQ`` list (``advection_diffusion2D.py:119-182``) as an operator:
    ``dQ[i] @ X`` (NumPy or CUDA tensor, (n,) or (n,k)) and ``dQ[i].tocsc()``.  Built on demand on the device."""

    def __init__(self, model, st, kind, payload):
        self.model, self.st, self.kind, self.payload = model, st, kind, payload
        self._slots = None
        n = model.engine.n
        self.shape = (n, n)

    def slots(self):
        if self._slots is None:
            m, eng = self.model, self.model.engine
            if self.kind == "q0":       # initial-field parameter: blockdiag(dQ0_j, 0, ..., 0)
                s = torch.zeros(eng.nslots * eng.n, dtype=F64, device=self.st["A9"].device)
                s0 = self.payload.slots().view(25, eng.Ns)
                s.view(eng.nslots, eng.n)[9:34, :eng.Ns] = s0
                self._slots = s
            else:
                self._slots = m._dQ_slots(self.st, self.kind, self.payload)
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
