"""Regular 2-D and 2-D x time meshes: host-side producers of the inputs of the CUDA hot path.

Mirrors the attribute/method surface of the reference's grid classes
(``grids/spatial2D_regular_mesh.py``, ``grids/spat2Dtemp_regular_mesh.py``): ``M N T Ns n hx hy V dt
Dv iDv bs bsH bsA Ne shape type sdim`` and ``setGrid extend getS n2e getIdx evalB evalBH evalAdv
advBound addCov addInt``.  Everything here is tiny host NumPy work done once per grid
(SURVEY.md section 2 row 4); the arrays it produces (cell volume, spline bases, selection
indices) are what the device kernels consume.

Cell index ``k = j*M + i`` (x fastest), space-time node ``k + t*Ns``
(``spat2Dtemp_regular_mesh.py:82-92``).
"""
from __future__ import annotations

import numpy as np
from scipy import sparse


def _bspline_deg2(pts: np.ndarray, lo: float, hi: float, nbs: int) -> np.ndarray:
    """Quadratic B-splines on ``nbs+5`` uniform knots padded by ``2*(hi-lo)/nbs`` with the two
    outermost pairs merged -> ``nbs`` functions (``spat2Dtemp_regular_mesh.py:157-198``).
    Cox-de Boor recursion, evaluated in the same order as the reference so values agree to the
    last bit."""
    pad = 2 * (hi - lo) / nbs
    kn = np.linspace(lo - pad, hi + pad, nbs + 5)
    nk = kn.size
    level = np.empty((pts.shape[0], nk - 1))
    for i in range(nk - 1):
        inside = (pts >= kn[i]) & (pts < kn[i + 1])
        if i == nk - 2:
            inside = inside | ((pts >= kn[i]) & (pts <= kn[i + 1]))
        level[:, i] = inside * 1.0
    for r in (1, 2):
        nxt = np.zeros((pts.shape[0], nk - r - 1))
        for i in range(nk - r - 1):
            nxt[:, i] = (pts - kn[i]) / (kn[i + r] - kn[i]) * level[:, i] \
                + (kn[i + r + 1] - pts) / (kn[i + r + 1] - kn[i + 1]) * level[:, i + 1]
        level = nxt
    first = (level[:, 0] + level[:, 1]).reshape(-1, 1)
    last = (level[:, -2] + level[:, -1]).reshape(-1, 1)
    return np.hstack([first, level[:, 2:-2], last])


class _RegularMesh:
    """Shared implementation; ``timed`` selects the space-time variant."""

    sdim = 2
    timed = False

    def __init__(self) -> None:
        self.type = "gridST" if self.timed else "gridS"
        self.meta = "Regular Mesh in 2D and time" if self.timed else "Regular Mesh in 2D"
        self.isExtended = False
        self.Ne = 0
        self.S = None
        self.cov = None
        self.inter = False
        self.scale = False
        self.Ae = None
        self._ramp = None
        self._plan_cache = {}
        # the reference's default mesh: 30 x 30 cells on [0,40]^2 (x10 time steps on [0,10])
        h = 40 / 30
        self.x = np.linspace(h / 2, 40 - h / 2, 30)
        self.y = np.linspace(h / 2, 40 - h / 2, 30)
        self.t = np.linspace(0.5, 9.5, 10) if self.timed else None
        self.setGrid()

    # ------------------------------------------------------------------ geometry
    def setGrid(self, x=None, y=None, t=None, extend=None, Nbs=3) -> None:
        self.x = self.x if x is None else np.asarray(x, dtype=np.float64)
        self.y = self.y if y is None else np.asarray(y, dtype=np.float64)
        self.M = self.x.shape[0]
        self.N = self.y.shape[0]
        self.A = self.x.max() - self.x.min()
        self.B = self.y.max() - self.y.min()
        self.hx = self.A / (self.M - 1)
        self.hy = self.B / (self.N - 1)
        self.V = self.hx * self.hy
        if self.timed:
            self.t = self.t if t is None else np.asarray(t, dtype=np.float64)
            self.T = self.t.shape[0]
            self.Tdur = self.t.max() - self.t.min()
            self.dt = self.Tdur / (self.T - 1)
        else:
            self.T = 1
        sx, sy = np.meshgrid(self.x, self.y)
        self.sx = sx.flatten()
        self.sy = sy.flatten()
        if self.timed:
            iy, it, ix = np.meshgrid(np.arange(self.N), np.arange(self.T), np.arange(self.M))
            self.ix, self.iy, self.it = ix.flatten(), iy.flatten(), it.flatten()
        self.isExtended = False
        self.Ne = 0
        self.Nbs = Nbs
        self.Nbs2 = Nbs ** 2
        self.Ns = self.M * self.N
        self.n = self.Ns * self.T
        self.S = None
        self.Ae = None
        self._ramp = None
        self._plan_cache = {}
        self.basisN()
        self.basisH()
        if self.timed:
            self.basisA()
        if extend is not None:
            self.extend(extend=extend)
        self.setDv()

    def extend(self, extend=1) -> None:
        """Grow the mesh by ``extend`` cells on every side (``spat2Dtemp_regular_mesh.py:102-117``)."""
        e = extend
        self.xe = np.hstack([np.linspace(self.x[0] - e * self.hx, self.x[0], e + 1)[:-1], self.x,
                             np.linspace(self.x[-1], self.x[-1] + e * self.hx, e + 1)[1:]])
        self.ye = np.hstack([np.linspace(self.y[0] - e * self.hy, self.y[0], e + 1)[:-1], self.y,
                             np.linspace(self.y[-1], self.y[-1] + e * self.hy, e + 1)[1:]])
        sxe, sye = np.meshgrid(self.xe, self.ye)
        self.sxe, self.sye = sxe.flatten(), sye.flatten()
        self.isExtended = True
        self.Ne = e
        self.Ns = (self.M + 2 * e) * (self.N + 2 * e)
        self.n = self.Ns * self.T
        self.S = None
        self._plan_cache = {}
        self.basisN()
        self.basisH()

    def setDv(self) -> None:
        self.Dv = self.V * sparse.eye(self.Ns)
        self.iDv = sparse.eye(self.Ns) / self.V

    @property
    def shape(self):
        if self.timed:
            return [self.M + 2 * self.Ne, self.N + 2 * self.Ne, self.T]
        return [self.M + 2 * self.Ne, self.N + 2 * self.Ne]

    # ------------------------------------------------------------------ selection matrix
    def n2e(self, idx):
        """original node index -> index in the extended mesh (``spat2Dtemp_regular_mesh.py:82-92``)."""
        a = np.asarray(idx)
        i = a % self.M
        j = (a // self.M) % self.N
        tt = a // (self.M * self.N)
        Me, Ne_ = self.M + 2 * self.Ne, self.N + 2 * self.Ne
        out = (i + self.Ne) + (j + self.Ne) * Me + tt * Me * Ne_
        return out if hasattr(idx, "__len__") else int(out)

    def getIdx(self, pos, extend=True):
        if self.timed:
            if extend:
                return (pos[0] + self.Ne) + (pos[1] + self.Ne) * (self.M + 2 * self.Ne) + pos[2] * self.Ns
            return pos[0] + pos[1] * self.M + pos[2] * self.M * self.N
        if extend:
            return (pos[0] + self.Ne) + (pos[1] + self.Ne) * (self.M + 2 * self.Ne)
        return pos[0] + pos[1] * self.M

    def setS(self) -> None:
        nobs = self.M * self.N * self.T
        cols = self.n2e(np.arange(nobs))
        S = sparse.csc_matrix((np.ones(nobs), (np.arange(nobs), cols)), shape=(nobs, int(np.prod(self.shape))))
        if self.cov is not None:
            if self.inter:
                c = self.cov / self.cov.max() if self.scale else self.cov
                S = sparse.bmat([[S, np.stack([np.ones(self.cov.shape[0]), c], axis=1)]])
            else:
                S = sparse.bmat([[S, self.cov.reshape(-1, 1)]])
        elif self.inter:
            S = sparse.bmat([[S, np.ones(nobs).reshape(-1, 1)]])
        self.S = S.tocsc()

    def getS(self, idxs=None) -> sparse.csc_matrix:
        if self.S is None:
            self.setS()
        if idxs is None:
            return self.S
        return self.S.tocsr()[np.asarray(idxs), :].tocsc()

    def obs_nodes(self, idxs=None) -> np.ndarray:
        """Extended-mesh node of every observation row of ``getS(idxs)`` (int64)."""
        base = self.n2e(np.arange(self.M * self.N * self.T))
        return base if idxs is None else base[np.asarray(idxs)]

    def addCov(self, cov, inter=True, scale=False) -> None:
        self.inter, self.cov = inter, cov
        if self.timed:
            self.scale = scale      # the spatial mesh of the reference never stores it (spatial2D_regular_mesh.py:55-58): no scaling there
        self.setS()

    def addInt(self) -> None:
        self.inter = True
        self.setS()

    # ------------------------------------------------------------------ spline bases
    def basis(self, dx=0.0, dy=0.0, d=2):
        tx = (self.sxe if self.isExtended else self.sx) + dx
        ty = (self.sye if self.isExtended else self.sy) + dy
        if dx != 0 or dy != 0:
            xlo, xhi = self.sx.min() - self.hx / 2, self.sx.max() + self.hx / 2
            ylo, yhi = self.sy.min() - self.hy / 2, self.sy.max() + self.hy / 2
        else:
            xlo, xhi, ylo, yhi = self.sx.min(), self.sx.max(), self.sy.min(), self.sy.max()
        return _bspline_deg2(tx, xlo, xhi, self.Nbs), _bspline_deg2(ty, ylo, yhi, self.Nbs)

    def _tensor(self, bx, by):
        # column i*Nbs + j = bx[..., j] * by[..., i]
        out = by[..., :, None] * bx[..., None, :]
        return out.reshape(out.shape[:-2] + (self.Nbs2,))

    def basisN(self) -> None:
        bx, by = self.basis()
        self.bs = self._tensor(bx, by)

    def _face_basis(self, order):
        shifts = {"W": (-self.hx / 2, 0.0), "E": (self.hx / 2, 0.0), "S": (0.0, -self.hy / 2), "N": (0.0, self.hy / 2)}
        bxs, bys = zip(*(self.basis(dx=shifts[f][0], dy=shifts[f][1]) for f in order))
        return self._tensor(np.stack(bxs, axis=1), np.stack(bys, axis=1))

    def basisH(self) -> None:
        self.bsH = self._face_basis("WESN")      # diffusion faces (AH_2D_b3.cpp:36-44)

    def basisA(self) -> None:
        self.bsA = self._face_basis("ENWS")      # advection faces (Aw_2D_b3.cpp:50-54)

    def evalB(self, par, bs=None, d=None):
        par = np.asarray(par, dtype=np.float64)
        if d is not None:
            par = np.zeros(par.shape)
            par[d] = 1
        return (self.bs if bs is None else bs) @ par

    def evalBH(self, par, bs=None, d=None):
        par = np.asarray(par, dtype=np.float64)
        if d is not None:
            par = np.zeros(par.shape)
            par[d] = 1
        return (self.bsH if bs is None else bs) @ par

    def evalAdv(self, par, bs=None, d=None):
        par = np.asarray(par, dtype=np.float64)
        if d is not None:
            par = np.zeros(par.shape)
            par[d] = 1
        bs = self.bsA if bs is None else bs
        n2 = self.Nbs ** 2
        res = np.stack([bs[:, 0, :] @ par[:n2], bs[:, 1, :] @ par[n2:], bs[:, 2, :] @ par[:n2], bs[:, 3, :] @ par[n2:]], axis=1)
        return self.advBound(res)

    # ------------------------------------------------------------------ extension ramp
    def _build_ramp(self):
        """Each extended cell takes ``scale * value(src)`` of one interior cell
        (``spat2Dtemp_regular_mesh.py:347-391``; the reference stores this as a dense matrix ``Ae``
        with a single entry per row)."""
        M, N, e = self.M, self.N, self.Ne
        Me, Ne_ = M + 2 * e, N + 2 * e
        src = np.zeros(Me * Ne_, dtype=np.int64)
        scl = np.zeros(Me * Ne_, dtype=np.float64)
        for j in range(Ne_):
            for i in range(Me):
                k = i + j * Me
                lo_i, hi_i, lo_j, hi_j = i < e, i >= M + e, j < e, j >= N + e
                if lo_i and lo_j:
                    src[k], scl[k] = 0, (j / e if i >= j else i / e)
                elif hi_i and hi_j:
                    src[k] = M * N - 1
                    scl[k] = (Ne_ - 1 - j) / e if (N - j <= M - i) else (Me - 1 - i) / e
                elif hi_i and lo_j:
                    src[k] = M - 1
                    scl[k] = (Me - 1 - i) / e if (Me - 1 - i <= j) else j / e
                elif lo_i and hi_j:
                    src[k] = M * (N - 1)
                    scl[k] = i / e if (i <= Ne_ - 1 - j) else (Ne_ - 1 - j) / e
                elif lo_i:
                    src[k], scl[k] = (j - e) * M, i / e
                elif hi_i:
                    src[k], scl[k] = M - 1 + (j - e) * M, (Me - 1 - i) / e
                elif lo_j:
                    src[k], scl[k] = i - e, j / e
                elif hi_j:
                    src[k], scl[k] = i - e + (N - 1) * M, (Ne_ - 1 - j) / e
                else:
                    src[k], scl[k] = (i - e) + (j - e) * M, 1.0
        self._ramp = (src, scl)

    def advBound(self, ww):
        if not self.isExtended:
            return ww
        if self._ramp is None:
            self._build_ramp()
        src, scl = self._ramp
        return scl[:, None] * ww[src]


    def assimilate_adv(self, we, wn) -> np.ndarray:
        """Cell-centred velocities (east, north components on the ``M x N`` cells, x fastest) -> the four face values
        ``(E, N, W, S)`` the covariate-driven advection classes take: mean of the two cells sharing the face, the cell's
        own value on the boundary, then the extension ramp (``spat2Dtemp_regular_mesh.py:277-344``; the reference builds
        four dense ``(MN)^2`` averaging matrices for this)."""
        M, N = self.M, self.N
        we = np.asarray(we, dtype=np.float64).reshape(N, M)
        wn = np.asarray(wn, dtype=np.float64).reshape(N, M)
        ww = np.zeros((M * N, 4))
        ww[:, 0] = ((we + np.concatenate([we[:, 1:], we[:, -1:]], axis=1)) / 2).reshape(-1)
        ww[:, 1] = ((wn + np.concatenate([wn[1:], wn[-1:]], axis=0)) / 2).reshape(-1)
        ww[:, 2] = ((we + np.concatenate([we[:, :1], we[:, :-1]], axis=1)) / 2).reshape(-1)
        ww[:, 3] = ((wn + np.concatenate([wn[:1], wn[:-1]], axis=0)) / 2).reshape(-1)
        return self.advBound(ww)

    def idx2pos(self, idx):
        return None          # a stub in the reference as well (spat2Dtemp_regular_mesh.py:70-71)

    def plot(self, value):
        return None          # a stub in the reference as well (:394-395)


class GridS(_RegularMesh):
    timed = False


class GridST(_RegularMesh):
    timed = True
