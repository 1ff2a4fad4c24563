"""Oracle-side regular mesh: the inputs of the reference's model classes.  TEST INFRASTRUCTURE ONLY.

Restates what ``grids/spatial2D_regular_mesh.py`` / ``grids/spat2Dtemp_regular_mesh.py`` of the reference compute for a
mesh WITHOUT boundary extension -- geometry (``:120-154``), selection matrix (``:35-58``), the quadratic B-spline bases
``bs / bsH / bsA`` (``:157-249``) and their evaluators (``:251-275``) -- so that ``bench.py --impl reference`` can build
its model without importing ``spdepy_b200``.  The splines are evaluated per axis on the M (or N) distinct coordinates
with the reference's own Cox-de Boor recursion and operation order, then combined, so the values are the reference's bit
for bit (``tests/test_oracle_grid.py`` compares with the fixtures recorded from the unmodified reference and with the
product's mesh class).  Extended meshes (``extend=``) are not needed by the bench workloads and are not restated here.
"""
from __future__ import annotations

import numpy as np
from scipy import sparse


def _bspline2_axis(pts, lo, hi, nbs):
    """``basis()`` of the reference along one axis (``spat2Dtemp_regular_mesh.py:174-186``): ``nbs + 5`` uniform knots
    padded by ``2 (hi - lo) / nbs``, degree-2 recursion, outer pairs merged."""
    kn = np.linspace(lo - 2 * (hi - lo) / nbs, hi + 2 * (hi - lo) / nbs, nbs + 5)
    B = [np.stack([((pts >= kn[i]) & (pts < kn[i + 1]) | ((pts >= kn[i]) & (pts <= kn[i + 1]) & (i == (kn.size - 2)))) * 1.0
                   for i in range(kn.size - 1)], axis=1)]
    for r in range(1, 3):
        B.append(np.zeros((pts.shape[0], kn.size - r - 1)))
        for i in range(kn.size - r - 1):
            B[r][:, i] = (pts - kn[i]) / (kn[i + r] - kn[i]) * B[r - 1][:, i] + (kn[i + r + 1] - pts) / (kn[i + r + 1] - kn[i + 1]) * B[r - 1][:, i + 1]
    return np.hstack([(B[2][:, 0] + B[2][:, 1]).reshape(-1, 1), B[2][:, 2:-2], (B[2][:, -2] + B[2][:, -1]).reshape(-1, 1)])


class OracleGrid:
    def __init__(self, x, y, t=None, Nbs=3):
        self.x, self.y = np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64)
        self.M, self.N = self.x.shape[0], self.y.shape[0]
        self.hx = (self.x.max() - self.x.min()) / (self.M - 1)
        self.hy = (self.y.max() - self.y.min()) / (self.N - 1)
        self.V = self.hx * self.hy
        self.timed = t is not None
        if self.timed:
            self.t = np.asarray(t, dtype=np.float64)
            self.T = self.t.shape[0]
            self.dt = (self.t.max() - self.t.min()) / (self.T - 1)
        else:
            self.T = 1
        self.Ns = self.M * self.N
        self.n = self.Ns * self.T
        self.Ne = 0
        self.Nbs, self.Nbs2 = Nbs, Nbs ** 2
        self.Dv = self.V * sparse.eye(self.Ns)
        self.iDv = sparse.eye(self.Ns) / self.V
        self._S = None
        kx, ky = np.arange(self.Ns) % self.M, np.arange(self.Ns) // self.M      # node k = j*M + i
        self._k = (kx, ky)
        self.bs = self._tensor(*self._axis_bases(0.0, 0.0))
        faces = {"W": (-1 / 2 * self.hx, 0.0), "E": (1 / 2 * self.hx, 0.0), "S": (0.0, -1 / 2 * self.hy), "N": (0.0, 1 / 2 * self.hy)}
        self.bsH = np.stack([self._tensor(*self._axis_bases(*faces[f])) for f in "WESN"], axis=1)     # :200-221
        if self.timed:
            self.bsA = np.stack([self._tensor(*self._axis_bases(*faces[f])) for f in "ENWS"], axis=1)  # :223-249

    @property
    def shape(self):
        return [self.M, self.N, self.T] if self.timed else [self.M, self.N]

    def _axis_bases(self, dx, dy):
        if dx != 0 or dy != 0:
            xlo, xhi = self.x.min() - self.hx / 2, self.x.max() + self.hx / 2
            ylo, yhi = self.y.min() - self.hy / 2, self.y.max() + self.hy / 2
        else:
            xlo, xhi, ylo, yhi = self.x.min(), self.x.max(), self.y.min(), self.y.max()
        return _bspline2_axis(self.x + dx, xlo, xhi, self.Nbs), _bspline2_axis(self.y + dy, ylo, yhi, self.Nbs)

    def _tensor(self, bx, by):
        kx, ky = self._k
        out = np.zeros((self.Ns, self.Nbs2))
        for i in range(self.Nbs):
            for j in range(self.Nbs):
                out[:, i * self.Nbs + j] = bx[kx, j] * by[ky, i]
        return out

    def evalB(self, par, bs=None, d=None):
        return (self.bs if bs is None else bs) @ np.asarray(par, dtype=np.float64)

    def evalBH(self, par, bs=None, d=None):
        return (self.bsH if bs is None else bs) @ np.asarray(par, dtype=np.float64)

    def evalAdv(self, par, bs=None, d=None):
        par = np.asarray(par, dtype=np.float64)
        bs = self.bsA if bs is None else bs
        n2 = self.Nbs2
        return np.stack([bs[:, 0, :] @ par[:n2], bs[:, 1, :] @ par[n2:], bs[:, 2, :] @ par[:n2], bs[:, 3, :] @ par[n2:]], axis=1)

    def getS(self, idxs=None):
        if self._S is None:
            self._S = sparse.eye(self.n, format="csc")
        if idxs is None:
            return self._S
        idxs = np.asarray(idxs)
        return sparse.csc_matrix((np.ones(idxs.size), (np.arange(idxs.size), idxs)), shape=(idxs.size, self.n))
