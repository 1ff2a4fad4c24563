"""CPU restatement of the reference's precision / likelihood / gradient algorithm.

TEST INFRASTRUCTURE ONLY.  Imported by tests/, by ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs and by ``__graft_entry__.smoke()`` as the *checker*; nothing under
``spdepy_b200/`` imports it and the product never falls back to it.

One spec-driven implementation restates what the reference spreads over its model classes
(scipy.sparse operations in the reference's own order, so ``Q`` comes out bit-identical):

  ``makeQ``    spatial   ``whittle_matern*2D.py:59-100``  (``Q = A^T iDv A``)
               space-time ``advection_diffusion2D.py:86-185`` and its siblings (block-tridiagonal
               ``Q``, explicit ``dQ/dtheta`` list)
  ``logLike``  ``advection_diffusion2D.py:187-223`` (identical text in every class)
  ``sample``   ``model.py:73-87``;  ``update`` ``model.py:120-127``

The stencils come from ``oracle/stencils.py`` (C restatement, bit-exact against the compiled
reference).  The sparse Cholesky of the reference lives in scikit-sparse 0.4.12 -> SuiteSparse
CHOLMOD (``poetry.lock:582-583``), absent from this image; its published supernodal algorithm is
restated in ``oracle/cpu_cholesky.py`` and a dense LAPACK stand-in is used for small ``n``.
PARITY STATUS: the reference's own tests hold no numbers for this path, so
the pins are (a) outputs of the unmodified reference Python run in the build container through
``oracle/ref_harness.py`` with the dense stand-in factoriser, committed as ``tests/golden/*.npz``
by ``oracle/make_golden.py``; (b) the analytic identities of SURVEY.md App. E.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
from scipy import sparse
from scipy import linalg as sla

import stencils as st


@dataclass(frozen=True)
class Spec:
    name: str            # reference ``type`` prefix, e.g. "advection-diffusion-2D"
    timed: bool
    kvar: bool           # spline kappa
    H: str               # "iso" | "aniso" | "ha"
    Hvar: bool           # per-face spline H
    w: str | None        # None | "const" | "var"
    aflav: str = "sum"   # "sum": Dv + Dv@Dk*dt - Ah*dt + Aw*dt ; "paren": Dv + (Dv@Dk - Ah + Aw)*dt
    qflav: str = "mul"   # "mul": 1/(dt*sigma)*Q ; "div": Q/(dt*sigma)


SPECS = {
    # spatial Whittle-Matern (spdes/__init__.py:3-24)
    "whittle-matern-isotropic-2D": Spec("whittle-matern-isotropic-2D", False, False, "iso", False, None),
    "whittle-matern-anisotropic-2D": Spec("whittle-matern-anisotropic-2D", False, False, "aniso", False, None),
    "whittle-matern-ha-2D": Spec("whittle-matern-ha-2D", False, False, "ha", False, None),
    "var-whittle-matern-anisotropic-2D": Spec("var-whittle-matern-anisotropic-2D", False, True, "aniso", True, None),
    # (assembled inline in the reference's logLike: var_whittle_matern2D.py:92-104, var_whittle_matern_ha2D.py:126-140)
    "var-whittle-matern-isotropic-2D": Spec("var-whittle-matern-isotropic-2D", False, True, "iso", True, None),
    "var-whittle-matern-ha-2D": Spec("var-whittle-matern-ha-2D", False, True, "ha", True, None),
    # advection-diffusion (spdes/__init__.py:25-35)
    "advection-diffusion-2D": Spec("advection-diffusion-2D", True, False, "aniso", False, "const"),
    "advection-idiffusion-2D": Spec("advection-idiffusion-2D", True, False, "iso", False, "const"),
    "advection-ha-diffusion-2D": Spec("advection-ha-diffusion-2D", True, False, "ha", False, "const"),
    # spatially varying families (spdes/__init__.py:36-46, 69-90)
    "advection-var-diffusion-2D": Spec("advection-var-diffusion-2D", True, True, "aniso", True, "const", "paren"),
    "advection-var-idiffusion-2D": Spec("advection-var-idiffusion-2D", True, True, "iso", True, "const", "paren"),
    "var-advection-diffusion-2D": Spec("var-advection-diffusion-2D", True, False, "aniso", False, "var", "paren"),
    "var-advection-idiffusion-2D": Spec("var-advection-idiffusion-2D", True, False, "iso", False, "var", "paren"),
    "cov-advection-diffusion-2D": Spec("cov-advection-diffusion-2D", True, False, "aniso", False, "cov"),
    "cov-advection-idiffusion-2D": Spec("cov-advection-idiffusion-2D", True, False, "iso", False, "cov"),
    "cov-advection-ha-diffusion-2D": Spec("cov-advection-ha-diffusion-2D", True, False, "ha", False, "cov"),
    "cov-advection-var-diffusion-2D": Spec("cov-advection-var-diffusion-2D", True, True, "aniso", True, "cov", "paren"),
    "cov-advection-var-idiffusion-2D": Spec("cov-advection-var-idiffusion-2D", True, True, "iso", True, "cov", "paren"),
    "cov-advection-var-ha-diffusion-2D": Spec("cov-advection-var-ha-diffusion-2D", True, True, "ha", True, "cov", "paren"),
    "advection-var-ha-diffusion-2D": Spec("advection-var-ha-diffusion-2D", True, True, "ha", True, "const", "sum"),
    "var-advection-ha-diffusion-2D": Spec("var-advection-ha-diffusion-2D", True, False, "ha", False, "var", "paren"),
    "var-advection-var-ha-diffusion-2D": Spec("var-advection-var-ha-diffusion-2D", True, True, "ha", True, "var", "paren", "div"),
    "var-advection-var-diffusion-2D": Spec("var-advection-var-diffusion-2D", True, True, "aniso", True, "var", "paren", "div"),
    "var-advection-var-idiffusion-2D": Spec("var-advection-var-idiffusion-2D", True, True, "iso", True, "var", "paren", "div"),
}


def n_own_params(spec: Spec, Np: int = 9) -> int:
    """Number of the model's own parameters *excluding* log tau."""
    nk = Np if spec.kvar else 1
    nh = {"iso": 1, "aniso": 3, "ha": 3}[spec.H] * (Np if spec.Hvar else 1)
    nw = {None: 0, "const": 2, "cov": 1, "var": 2 * Np}[spec.w]
    return nk + nh + nw + (1 if spec.timed else 0)


# ---------------------------------------------------------------------------------------------
# factor stand-ins

class DenseFactor:
    """``L L^T = P A P^T`` by LAPACK; same methods as ``sksparse.cholmod.Factor`` used at
    ``advection_diffusion2D.py:194-202`` / ``model.py:80,126``.  ``perm`` is new -> old."""

    def __init__(self, A, perm=None):
        A = sparse.csc_matrix(A)
        n = A.shape[0]
        self.perm = np.arange(n) if perm is None else np.asarray(perm, dtype=np.int64)
        Al = sparse.tril(A).toarray()            # CHOLMOD reads the lower triangle only
        Ad = Al + np.tril(Al, -1).T
        self.L = np.linalg.cholesky(Ad[np.ix_(self.perm, self.perm)])

    def P(self):
        return self.perm.copy()

    def logdet(self):
        return 2.0 * np.log(np.diag(self.L)).sum()

    def apply_Pt(self, x):
        out = np.empty_like(np.asarray(x, dtype=np.float64))
        out[self.perm] = x
        return out

    def solve_Lt(self, b, use_LDLt_decomposition=False):
        return sla.solve_triangular(self.L, np.asarray(b, dtype=np.float64), lower=True, trans="T")

    def solve_A(self, b):
        b = np.asarray(b.toarray() if sparse.issparse(b) else b, dtype=np.float64)
        y = sla.solve_triangular(self.L, b[self.perm], lower=True)
        return self.apply_Pt(sla.solve_triangular(self.L, y, lower=True, trans="T"))


_factor_impl = DenseFactor
_perm_provider = None


def set_factor(impl=None, perm_provider=None):
    """Choose the factoriser used by the oracle (``DenseFactor`` or ``cpu_cholesky.SupernodalFactor``)
    and, optionally, a callable ``n -> perm`` so samples use the build's permutation."""
    global _factor_impl, _perm_provider
    _factor_impl = DenseFactor if impl is None else impl
    _perm_provider = perm_provider


def cholesky(A):
    perm = None if _perm_provider is None else _perm_provider(A.shape[0])
    return _factor_impl(A, perm=perm)


# ---------------------------------------------------------------------------------------------

class OracleSPDE:
    """Restatement of one reference model class; ``grid`` is any object with the reference grid
    attributes (``M N Ns n T hx hy V dt Dv iDv bs bsH bsA shape evalB evalBH evalAdv getS``)."""

    def __init__(self, spec: Spec | str, grid, mod0: "OracleSPDE | None" = None, par=None, bc: int = 3, ww=None):
        self.ww = ww      # cov-advection: supplied face velocities (cov_advection_diffusion2D.py:21)
        self.spec = SPECS[spec] if isinstance(spec, str) else spec
        self.grid = grid
        self.bc = bc
        self.mod0 = mod0
        self.Np = grid.Nbs2
        self.type = "%s-bc%d" % (self.spec.name, bc)
        self.Q = None
        self.Q_fac = None
        self.data = None
        self.r = None
        self.S = None
        self.par = None
        if self.spec.timed:
            assert mod0 is not None
        if par is not None:
            self.setQ(par)

    # --- stencil wrappers (advection_diffusion2D.py:226-260, var_advection_var_diffusion2D.py:239-276)
    def Ah(self, Hs):
        M, N = self.grid.shape[0], self.grid.shape[1]
        Hs = np.array(Hs, dtype="float64")
        if Hs.ndim == 2:
            tri = st.oracle_ah_const(M, N, Hs, self.grid.hx, self.grid.hy, self.bc)
        else:
            tri = st.oracle_ah_face(M, N, Hs, self.grid.hx, self.grid.hy, self.bc)
        return st.to_csc(tri, M * N)

    def Aw(self, ws, dws=None, diff=3, nan_to_zero=True):
        M, N = self.grid.shape[0], self.grid.shape[1]
        ws = np.array(ws, dtype="float64")
        if ws.ndim == 2 and not nan_to_zero:      # cov_advection_diffusion2D.py:225-240: no NaN filter
            return st.to_csc(st.oracle_aw_face(M, N, ws, None, self.grid.hx, self.grid.hy, diff, self.bc), M * N)
        if ws.ndim == 1:
            return st.to_csc(st.oracle_aw_const(M, N, ws, self.grid.hx, self.grid.hy, diff, self.bc), M * N)
        tri = st.oracle_aw_face(M, N, ws, dws, self.grid.hx, self.grid.hy, diff, self.bc)
        return st.to_csc(tri, M * N, nan_to_zero=True)

    # --- parameter -> fields and derivative directions
    def _split(self, par):
        sp, Np = self.spec, self.Np
        nk = Np if sp.kvar else 1
        nh1 = Np if sp.Hvar else 1
        o = 0
        out = {"kappa": par[o:o + nk]}
        o += nk
        out["gamma"] = par[o:o + nh1]
        o += nh1
        if sp.H != "iso":
            out["vx"] = par[o:o + nh1]
            out["vy"] = par[o + nh1:o + 2 * nh1]
            o += 2 * nh1
        if sp.w == "const":
            out["w"] = par[o:o + 2]
            o += 2
        elif sp.w == "var":
            out["w"] = par[o:o + 2 * Np]
            o += 2 * Np
        elif sp.w == "cov":
            out["w"] = par[o:o + 1]
            o += 1
        if sp.timed:
            out["sigma"] = par[o]
            o += 1
        out["n_own"] = o
        return out

    def _H(self, p):
        """H and the list of (parameter-name, dH) directions in the reference's parameter order."""
        sp, g = self.spec, self.grid
        dirs = []
        if not sp.Hvar:
            gamma = np.exp(p["gamma"][0])
            if sp.H == "iso":
                Hs = gamma * np.eye(2)
                dirs.append(("gamma", gamma * np.eye(2)))
            elif sp.H == "aniso":
                vv = np.array([p["vx"][0], p["vy"][0]])
                Hs = gamma * np.eye(2) + vv[:, np.newaxis] * vv[np.newaxis, :]
                dirs.append(("gamma", gamma * np.eye(2)))
                for e in (np.array([1.0, 0.0]), np.array([0.0, 1.0])):
                    dirs.append(("v", e[:, np.newaxis] * vv[np.newaxis, :] + vv[:, np.newaxis] * e[np.newaxis, :]))
            else:   # half-angle, whittle_matern_ha2D.py:73-106
                vx, vy = p["vx"][0], p["vy"][0]
                aV = np.sqrt(vx ** 2 + vy ** 2)
                mV = np.array([[vx, vy], [vy, -vx]])
                ch = (np.exp(aV) + np.exp(-aV)) / 2
                sh = (np.exp(aV) - np.exp(-aV)) / 2
                Hs = gamma * (ch * np.eye(2) + sh / aV * mV)
                dirs.append(("gamma", Hs))
                dirs.append(("v", gamma / aV * (vx * sh * np.eye(2) + vx / aV * (ch - sh / aV) * mV + sh * np.array([[1, 0], [0, -1]]))))
                dirs.append(("v", gamma / aV * (vy * sh * np.eye(2) + vy / aV * (ch - sh / aV) * mV + sh * np.array([[0, 1], [1, 0]]))))
            return Hs, dirs
        gamma = np.exp(g.evalBH(par=p["gamma"]))
        eye = np.eye(2)
        if sp.H == "ha":        # advection_var_ha_diffusion2D.py:104-113,145-175
            vx, vy = g.evalBH(p["vx"]), g.evalBH(p["vy"])
            aV = np.sqrt(vx ** 2 + vy ** 2)
            mV = np.array([[vx, vy], [vy, -vx]]).T.swapaxes(0, 1)
            ch = (np.exp(aV) + np.exp(-aV)) / 2
            sh = (np.exp(aV) - np.exp(-aV)) / 2
            nx = np.newaxis
            Hs = (gamma * ch)[:, :, nx, nx] * eye + (gamma * sh / aV)[:, :, nx, nx] * mV
            for i in range(self.Np):
                dg = g.bsH[:, :, i] * gamma
                dirs.append(("gamma", (dg * ch)[:, :, nx, nx] * eye + (dg * sh / aV)[:, :, nx, nx] * mV))
            for comp in (0, 1):
                for i in range(self.Np):
                    dpar = np.zeros(self.Np)
                    dpar[i] = 1
                    dv = g.evalBH(par=dpar)
                    z = vx * 0
                    dmV = (np.array([[dv, z], [z, -dv]]) if comp == 0 else np.array([[z, dv], [dv, -z]])).T.swapaxes(0, 1)
                    vc = vx if comp == 0 else vy
                    dirs.append(("v", (gamma * dv * vc / aV)[:, :, nx, nx] * (sh[:, :, nx, nx] * eye + ((ch - sh / aV) / aV)[:, :, nx, nx] * mV)
                                 + (gamma * sh / aV)[:, :, nx, nx] * dmV))
            return Hs, dirs
        if sp.H == "iso":
            Hs = eye * (np.stack([gamma, gamma], axis=2))[:, :, :, np.newaxis]
        else:
            vx, vy = g.evalBH(p["vx"]), g.evalBH(p["vy"])
            vv = np.stack([vx, vy], axis=2)
            Hs = (eye * (np.stack([gamma, gamma], axis=2))[:, :, :, np.newaxis]) + vv[:, :, :, np.newaxis] * vv[:, :, np.newaxis, :]
        for i in range(self.Np):
            dg = g.bsH[:, :, i] * gamma
            dirs.append(("gamma", eye * (np.stack([dg, dg], axis=2)[:, :, :, np.newaxis])))
        if sp.H == "aniso":
            zero = g.evalBH(par=np.zeros(self.Np))
            for comp in (0, 1):
                for i in range(self.Np):
                    dpar = np.zeros(self.Np)
                    dpar[i] = 1
                    b = g.evalBH(par=dpar)
                    dv = np.stack([b, zero], axis=2) if comp == 0 else np.stack([zero, b], axis=2)
                    dirs.append(("v", vv[:, :, :, np.newaxis] * dv[:, :, np.newaxis, :] + dv[:, :, :, np.newaxis] * vv[:, :, np.newaxis, :]))
        return Hs, dirs

    # --- spatial model
    def _makeQ_spatial(self, par, grad):
        g, sp = self.grid, self.spec
        Dv, iDv, Ns = g.Dv, g.iDv, g.Ns
        p = self._split(par)
        if sp.kvar:
            kappa = np.exp(g.evalB(par=p["kappa"]))
            Dk = sparse.diags(kappa).tocsc()
        else:
            kappa = np.exp(p["kappa"][0])
            Dk = kappa * sparse.eye(Ns)
        Hs, dirs = self._H(p)
        A = Dv @ Dk - self.Ah(Hs)
        Q = A.transpose() @ iDv @ A
        Q_fac = cholesky(Q)
        if not grad:
            return Q, Q_fac, None
        dQ = []
        if sp.kvar:
            for i in range(self.Np):
                dA = Dv @ sparse.diags(g.bs[:, i] * kappa)
                dQ.append((dA.transpose() @ iDv @ A + A.transpose() @ iDv @ dA).tocsc())
        else:
            dA = (Dv @ Dk).tocsc()
            dQ.append(dA.T @ iDv @ A + A.T @ iDv @ dA)
        for _, dH in dirs:
            dA = -self.Ah(dH)
            dQ.append((dA.transpose() @ iDv @ A + A.transpose() @ iDv @ dA).tocsc())
        return Q, Q_fac, dQ

    # --- space-time model (advection_diffusion2D.py:86-185)
    def _st_blocks(self, first, mid, last, up, low, Ns, T):
        Z = sparse.csc_matrix
        Q = sparse.bmat([[first, up, Z((Ns, (T - 2) * Ns))]])
        for t in range(T - 2):
            Q = sparse.bmat([[Q], [sparse.bmat([[Z((Ns, t * Ns)), low, mid, up, Z((Ns, (T - 3 - t) * Ns))]])]])
        Q = sparse.bmat([[Q], [sparse.bmat([[Z((Ns, (T - 2) * Ns)), low, last]])]])
        return Q

    def _makeQ_timed(self, par, grad):
        g, sp = self.grid, self.spec
        dt, T, Ns, Dv, iDv = g.dt, g.T, g.Ns, g.Dv, g.iDv
        p = self._split(par)
        if sp.kvar:
            kappa = np.exp(g.evalB(p["kappa"]))
            Dk = sparse.diags(kappa).tocsc()
        else:
            kappa = np.exp(p["kappa"][0])
            Dk = kappa * sparse.eye(Ns)
        Hs, dirs = self._H(p)
        if sp.w == "cov":
            ws = p["w"][0] * self.ww
            Aw_val = self.Aw(ws, nan_to_zero=False)
        else:
            ws = p["w"] if sp.w == "const" else g.evalAdv(p["w"])
            Aw_val = self.Aw(ws)
        sigma = np.exp(p["sigma"])
        As = Dv @ Dk
        Qs = As.transpose() @ iDv @ As
        if sp.aflav == "sum":
            A = Dv + Dv @ Dk * dt - self.Ah(Hs) * dt + Aw_val * dt
        else:
            A = Dv + (Dv @ Dk - self.Ah(Hs) + Aw_val) * dt
        n_own = p["n_own"]
        dQ0 = None
        if par.size > n_own + 1:
            Q0, _, dQ0 = self.mod0.makeQ(par=par[n_own:], grad=grad)
        else:
            if self.mod0.Q is None:
                self.mod0.setQ()
            Q0 = self.mod0.Q
        Q = self._st_blocks(sigma * dt * Q0 + Qs, A.T @ iDv @ Qs @ iDv @ A + Qs, A.T @ iDv @ Qs @ iDv @ A,
                            -Qs @ iDv @ A, -A.T @ iDv @ Qs, Ns, T)
        Q = 1 / (dt * sigma) * Q.tocsc() if sp.qflav == "mul" else Q.tocsc() / (dt * sigma)
        Q_fac = cholesky(Q)
        if not grad:
            return Q, Q_fac, None

        def scale(tdQ, sign=1.0):
            if sp.qflav == "mul":
                return (sign * 1 / (dt * sigma) * tdQ).tocsc()
            return (sign * tdQ / (dt * sigma)).tocsc()

        def dQ_of_dA(dA):
            low = -dA.T @ iDv @ Qs
            up = -Qs @ iDv @ dA
            mid = dA.T @ iDv @ Qs @ iDv @ A + A.T @ iDv @ Qs @ iDv @ dA
            return scale(self._st_blocks(sparse.csc_matrix((Ns, Ns)), mid, mid, up, low, Ns, T))

        dQ = []
        # log kappa (advection_diffusion2D.py:121-129, var_advection_var_diffusion2D.py:120-129)
        kdirs = [(Dv @ Dk * dt).tocsc()] if not sp.kvar else None
        for i in range(self.Np if sp.kvar else 1):
            if sp.kvar:
                dA = dt * Dv @ (sparse.diags(g.bs[:, i] * kappa).tocsc())
                dAs = Dv @ (sparse.diags(g.bs[:, i] * kappa).tocsc())
            else:
                dA = kdirs[0]
                dAs = Dv @ Dk
            dQs = dAs.T @ iDv @ As + As.T @ iDv @ dAs
            low = -dA.T @ iDv @ Qs - A.T @ iDv @ dQs
            up = -dQs @ iDv @ A - Qs @ iDv @ dA
            core = dA.T @ iDv @ Qs @ iDv @ A + A.T @ iDv @ dQs @ iDv @ A + A.T @ iDv @ Qs @ iDv @ dA
            dQ.append(scale(self._st_blocks(dQs, core + dQs, core, up, low, Ns, T)))
        # diffusion parameters
        for _, dH in dirs:
            dQ.append(dQ_of_dA(-self.Ah(dH) * dt))
        # advection parameters
        if sp.w == "const":
            for d in (1, 2):
                dQ.append(dQ_of_dA(self.Aw(ws, diff=d) * dt))
        elif sp.w == "cov":
            dQ.append(dQ_of_dA(self.Aw(self.ww, nan_to_zero=False) * dt))
        elif sp.w == "var":
            for i in range(2 * self.Np):
                dpar = np.zeros(self.Np * 2)
                dpar[i] = 1
                dws = g.evalAdv(dpar)
                dQ.append(dQ_of_dA(self.Aw(ws, dws, diff=1 if i < self.Np else 2) * dt))
        # log sigma (advection_diffusion2D.py:170-175)
        tdQ = self._st_blocks(Qs, A.T @ iDv @ Qs @ iDv @ A + Qs, A.T @ iDv @ Qs @ iDv @ A, -Qs @ iDv @ A, -A.T @ iDv @ Qs, Ns, T)
        dQ.append(scale(tdQ, -1.0))
        # initial-field parameters (advection_diffusion2D.py:176-182)
        if dQ0 is not None:
            for dq in dQ0:
                dQ.append(sparse.block_diag([dq, sparse.csc_matrix(((T - 1) * Ns, (T - 1) * Ns))]).tocsc())
        return Q, Q_fac, dQ

    # --- the reference's public methods
    def makeQ(self, par, grad=True):
        par = np.asarray(par, dtype="float64")
        return self._makeQ_timed(par, grad) if self.spec.timed else self._makeQ_spatial(par, grad)

    def setQ(self, par=None):
        if par is None:
            par = self.par
        self.par = np.asarray(par, dtype="float64")
        self.tau = self.par[-1]
        n_own = self._split(self.par)["n_own"]
        if self.spec.timed and self.par.size > n_own + 1:
            self.mod0.setQ(self.par[n_own:])
        self.Q, self.Q_fac, _ = self.makeQ(self.par, grad=False)
        self.S = self.grid.getS()

    def initFit(self, data, idx=None):
        self.data = data
        self.r = data.shape[1] if data.ndim == 2 else 1
        self.S = self.grid.getS(idxs=idx)

    def logLike(self, par, nh1=100, grad=True, probes=None):
        """``advection_diffusion2D.py:187-223``.  ``probes`` replaces the draw from the global
        legacy RNG at line 200 (the caller seeds or injects)."""
        par = np.asarray(par, dtype="float64")
        data = self.data
        tau = np.exp(par[-1])
        S = self.S
        Q, Q_fac, dQ = self.makeQ(par=par, grad=grad)
        Q_c = Q + S.T @ S * tau
        Q_c_fac = cholesky(Q_c)
        mu_c = Q_c_fac.solve_A(S.T @ data * tau)
        if self.r == 1:
            data = data.reshape(-1, 1)
            mu_c = mu_c.reshape(-1, 1)
        nobs = S.shape[0]
        like = 1 / 2 * Q_fac.logdet() * self.r + nobs * self.r * np.log(tau) / 2 - 1 / 2 * Q_c_fac.logdet() * self.r \
            - 1 / 2 * (mu_c * (Q @ mu_c)).sum() - tau / 2 * ((data - S @ mu_c) ** 2).sum()
        self.last = {"mu_c": mu_c, "logdetQ": Q_fac.logdet(), "logdetQc": Q_c_fac.logdet(), "Q": Q}
        if not grad:
            return -like / (nobs * self.r)
        if probes is None:
            vtmp = (2 * np.random.randint(1, 3, self.grid.n * nh1) - 3).reshape(self.grid.n, nh1)
        else:
            vtmp = np.asarray(probes)
            nh1 = vtmp.shape[1]
        TrQ = Q_fac.solve_A(vtmp)
        TrQc = Q_c_fac.solve_A(vtmp)
        g_par = np.zeros(par.size)
        for i in range(par.size - 1):
            dQmu_c = dQ[i] @ mu_c
            g_par[i] = (1 / 2 * ((TrQ - TrQc) * (dQ[i] @ vtmp)).sum() * self.r / nh1 - 1 / 2 * (mu_c * dQmu_c).sum())
        g_par[-1] = nobs * self.r / 2 - 1 / 2 * (TrQc * (S.T @ S * tau @ vtmp)).sum() * self.r / nh1 \
            - tau / 2 * ((data - S @ mu_c) ** 2).sum()
        return -like / (nobs * self.r), -g_par / (nobs * self.r)

    def logLike_exact(self, par):
        """Same objective with the Hutchinson traces replaced by exact ``tr(Q^-1 dQ_i)`` from a dense
        inverse (the expectation of ``advection_diffusion2D.py:200-207``); small grids only."""
        par = np.asarray(par, dtype="float64")
        data = self.data
        tau = np.exp(par[-1])
        S = self.S
        Q, Q_fac, dQ = self.makeQ(par=par, grad=True)
        Q_c = (Q + S.T @ S * tau).tocsc()
        Q_c_fac = cholesky(Q_c)
        mu_c = Q_c_fac.solve_A(S.T @ data * tau)
        if self.r == 1:
            data = data.reshape(-1, 1)
            mu_c = mu_c.reshape(-1, 1)
        nobs = S.shape[0]
        like = 1 / 2 * Q_fac.logdet() * self.r + nobs * self.r * np.log(tau) / 2 - 1 / 2 * Q_c_fac.logdet() * self.r \
            - 1 / 2 * (mu_c * (Q @ mu_c)).sum() - tau / 2 * ((data - S @ mu_c) ** 2).sum()
        Qd = sparse.tril(Q).toarray()
        Qd = Qd + np.tril(Qd, -1).T
        Qcd = sparse.tril(Q_c).toarray()
        Qcd = Qcd + np.tril(Qcd, -1).T
        Zi, Zci = np.linalg.inv(Qd), np.linalg.inv(Qcd)
        D = Zi - Zci
        g_par = np.zeros(par.size)
        for i in range(par.size - 1):
            g_par[i] = 1 / 2 * self.r * (dQ[i].multiply(D.T)).sum() - 1 / 2 * (mu_c * (dQ[i] @ mu_c)).sum()
        StS = (S.T @ S * tau).tocsc()
        g_par[-1] = nobs * self.r / 2 - 1 / 2 * self.r * (StS.multiply(Zci.T)).sum() - tau / 2 * ((data - S @ mu_c) ** 2).sum()
        return -like / (nobs * self.r), -g_par / (nobs * self.r)


class OracleSeparable(OracleSPDE):
    """Restatement of the separable classes ``Q = kron(Qt, Qs)``, ``Qs`` a spatially varying Whittle-Matern precision:

    * ``variant="ani"``: ``SeperableSpatialTemporal2D`` (``seperable_spatial_temporal2D.py:64-122, 186-211``), ``Qt`` AR(1)
      in ``rho``; ``par = [kappa x9, gamma x9, vx x9, vy x9, log rho, log tau]``;
    * ``variant="ha"``: ``SeperableSpatialTemporalHa2D`` (``seperable_spatial_temporal_ha2D.py:70-138, 202-226``), half-angle
      ``H``, ``Qt = sigma tridiag(-a, 1 + a^2, -a)`` with ``sigma`` alone in the two corners;
      ``par = [kappa x9, gamma x9, vx x9, vy x9, a, log sigma, log tau]``, ``dQ`` ends with ``kron(dQt/da, Qs)`` and ``Q``;
    * ``variant="iso"``: ``SeperableSpatialTemporalIDiffusion2D`` (``seperable_spatial_temporal_idiffusion2D.py:64-102,
      166-190``), isotropic ``H``; ``par = [kappa x9, gamma x9, a, log sigma, log tau]``.

    The ``ha`` / ``iso`` classes fill ``Qt`` with a hard-coded ``range(10)`` (``..._ha2D.py:205,216``): they are defined for
    ``T = 10`` only, which is what this restatement insists on.  ``logLike`` / ``logLike_exact`` are inherited (the
    reference's ``logLike`` body is the common one, ``:125-170``)."""

    KEYS = {"ani": ("var-whittle-matern-anisotropic-2D", "seperable-spatial-temporal-ani-2D-bc%d", 36),
            "ha": ("var-whittle-matern-ha-2D", "seperable-spatial-temporal-ha-2D-bc%d", 36),
            "iso": ("var-whittle-matern-isotropic-2D", "seperable-spatial-temporal-2D-bc%d", 18)}

    def __init__(self, grid, par=None, bc: int = 3, variant: str = "ani"):
        key, typ, self.nsp = self.KEYS[variant]
        super().__init__(key, grid, None, None, bc)
        self.variant = variant
        self.type = typ % bc
        if variant != "ani" and grid.T != 10:
            raise ValueError("the reference builds Qt of this class for T = 10 only")
        if par is not None:
            self.setQ(par)

    @staticmethod
    def makeQt(rho, T, diff=0):
        res = np.zeros((T, T))
        for i in range(T):
            if diff == 1:
                res[i, i] = (2 if i == 0 or i == T - 1 else 4) * rho ** 2 / (1 - rho ** 2) ** 2
                off = -rho * (1 + rho ** 2) / (1 - rho ** 2) ** 2
            else:
                res[i, i] = 1 / (1 - rho ** 2) if i == 0 or i == T - 1 else (1 + rho ** 2) / (1 - rho ** 2)
                off = -rho / (1 - rho ** 2)
            if i > 0:
                res[i, i - 1] = off
            if i < T - 1:
                res[i, i + 1] = off
        return sparse.csc_matrix(res)

    @staticmethod
    def makeQt_a(a, sigma, T, diff=0):
        """``makeQt(a, sigma)`` of the ha / idiffusion classes; ``diff=1`` is the derivative with respect to ``a``."""
        res = np.zeros((T, T))
        for i in range(T):
            ends = i == 0 or i == T - 1
            if diff == 1:
                res[i, i] = 0 if ends else 2 * a * sigma
                off = -sigma
            else:
                res[i, i] = sigma if ends else (1 + a ** 2) * sigma
                off = -a * sigma
            if i > 0:
                res[i, i - 1] = off
            if i < T - 1:
                res[i, i + 1] = off
        return sparse.csc_matrix(res)

    def makeQ(self, par, grad=True):
        par = np.asarray(par, dtype="float64")
        T, k = self.grid.T, self.nsp
        Qs, _, dQs = self._makeQ_spatial(np.hstack([par[:k], par[-1]]), grad)
        if self.variant == "ani":
            rho = np.exp(par[k])
            Qt = self.makeQt(rho, T)
        else:
            apar, sigma = par[k], np.exp(par[k + 1])
            Qt = self.makeQt_a(apar, sigma, T)
        Q = sparse.kron(Qt, Qs).tocsc()
        Q_fac = cholesky(Q)
        if not grad:
            return Q, Q_fac, None
        dQ = [sparse.kron(Qt, d).tocsc() for d in dQs]
        if self.variant == "ani":
            dQ.append(sparse.kron(self.makeQt(rho, T, diff=1), Qs).tocsc())
        else:
            dQ.append(sparse.kron(self.makeQt_a(apar, sigma, T, diff=1), Qs).tocsc())
            dQ.append(Q)          # log sigma: Qt is proportional to sigma
        return Q, Q_fac, dQ

    def setQ(self, par=None):
        if par is None:
            par = self.par
        self.par = np.asarray(par, dtype="float64")
        self.tau = self.par[-1]
        self.Q, self.Q_fac, _ = self.makeQ(self.par, grad=False)
        self.S = self.grid.getS()


def sample(Q, S_full, n=1, seed=0, mu=None, tau=None, simple=True, nprod=None):
    """``Model.sample`` (``model.py:73-87``): ``S (P^T L^-T z + mu)`` (+ ``S[:, :N] z[:N]/sqrt(tau)``)."""
    N = Q.shape[0]
    z = np.random.default_rng(seed).normal(size=N * n).reshape(N, n)
    fac = cholesky(Q)
    mu = np.zeros(N) if mu is None else mu
    data = S_full @ (fac.apply_Pt(fac.solve_Lt(z, use_LDLt_decomposition=False)) + mu[:, np.newaxis])
    if not simple:
        data += S_full[:, :nprod] @ z[:nprod] * 1 / np.sqrt(tau)
    return data


def update(Q, mu, S, y, tau):
    """``Model.update`` (``model.py:120-127``)."""
    Q = Q + S.transpose() @ S * tau
    fac = cholesky(Q)
    tmp = fac.solve_A(S.T @ (y - S @ mu)) * tau
    return Q, mu + tmp
