"""CPU supernodal multifrontal Cholesky -- the oracle's stand-in for CHOLMOD.  TEST INFRASTRUCTURE ONLY.

The reference factorises through scikit-sparse 0.4.12 -> SuiteSparse CHOLMOD (``poetry.lock:582-583``;
call sites ``advection_diffusion2D.py:117,193``, ``model.py:79,125``), a system library that is absent
from ``/root/reference`` and from this image.  This module restates CHOLMOD's published supernodal
algorithm (Chen, Davis, Hager, Rajamanickam, ACM TOMS 35(3), 2008: supernodal ``L L^T`` with dense
``potrf / trsm / syrk`` on the supernodes) in multifrontal form with NumPy / LAPACK dense kernels.
The fill-reducing permutation and the supernode partition are taken from the build's own symbolic
analysis (``spde_plan_perm`` / ``spde_plan_supernodes``, host code), because ``P^T L^-T z`` samples
are only comparable under the same ``P`` (SURVEY.md finding 3).

It is used (a) as the independent numeric check of the CUDA factorisation at sizes where a dense
factor is too big, and (b) as the CPU baseline that ``bench.py`` times on the GPU box's host cores.
PARITY STATUS: no reference test pins numbers at the CHOLMOD boundary; this stand-in
is itself checked against dense LAPACK in tests/test_cpu_cholesky.py.
"""
from __future__ import annotations

import numpy as np
from scipy import linalg as sla
from scipy import sparse


class SupernodalFactor:
    """``L L^T = P A P^T``; same method surface as ``sksparse.cholmod.Factor`` as used by the
    reference: ``logdet, solve_A, solve_Lt, solve_L, apply_P, apply_Pt, P``."""

    def __init__(self, A, perm=None, plan=None):
        if plan is None:
            raise ValueError("SupernodalFactor needs the build's symbolic plan (PlanHandle)")
        A = sparse.csc_matrix(A)
        n = A.shape[0]
        self.n = n
        self.perm = plan.perm.astype(np.int64)
        first, rowptr, rows, parent = plan.supernodes()
        self.first, self.rowptr, self.rows, self.parent = first, rowptr, rows.astype(np.int64), parent
        ns = first.size - 1
        # CHOLMOD reads the lower triangle of A; symmetric permutation of it
        Al = sparse.tril(A, format="coo")
        ip = np.empty(n, np.int64)
        ip[self.perm] = np.arange(n)
        r, c = ip[Al.row], ip[Al.col]
        lo = np.minimum(r, c)
        hi = np.maximum(r, c)
        Ap = sparse.csc_matrix((Al.data, (hi, lo)), shape=(n, n))
        Ap.sort_indices()
        indptr, indices, data = Ap.indptr, Ap.indices, Ap.data
        self.L11 = [None] * ns
        self.L21 = [None] * ns
        upd = [None] * ns
        kids = [[] for _ in range(ns)]
        for s in range(ns):
            if parent[s] >= 0:
                kids[parent[s]].append(s)
        for s in range(ns):
            f, l = int(first[s]), int(first[s + 1])
            nc = l - f
            below = self.rows[rowptr[s]:rowptr[s + 1]]
            nr = below.size
            m = nc + nr
            F = np.zeros((m, m))
            # original entries of the pivot columns
            lo_p, hi_p = indptr[f], indptr[l]
            ri = indices[lo_p:hi_p]
            cj = np.repeat(np.arange(nc), np.diff(indptr[f:l + 1]))
            pos = np.where(ri < l, ri - f, nc + np.searchsorted(below, ri))
            F[pos, cj] = data[lo_p:hi_p]
            # extend-add of the children's update matrices
            for c_ in kids[s]:
                cb = self.rows[rowptr[c_]:rowptr[c_ + 1]]
                rel = np.where(cb < l, cb - f, nc + np.searchsorted(below, cb))
                F[np.ix_(rel, rel)] += upd[c_]
                upd[c_] = None
            L11 = np.linalg.cholesky(F[:nc, :nc] + np.tril(F[:nc, :nc], -1).T)
            self.L11[s] = L11
            if nr:
                L21 = sla.solve_triangular(L11, F[nc:, :nc].T, lower=True).T
                self.L21[s] = L21
                U = F[nc:, nc:]
                U = U + np.tril(U, -1).T
                upd[s] = U - L21 @ L21.T
            else:
                self.L21[s] = np.zeros((0, nc))

    def P(self):
        return self.perm.copy()

    def logdet(self):
        return 2.0 * sum(np.log(np.diag(L)).sum() for L in self.L11)

    def apply_P(self, x):
        return np.asarray(x)[self.perm]

    def apply_Pt(self, x):
        x = np.asarray(x, dtype=np.float64)
        out = np.empty_like(x)
        out[self.perm] = x
        return out

    def solve_L(self, b, use_LDLt_decomposition=False):
        x = np.array(b, dtype=np.float64)
        one = x.ndim == 1
        x = x.reshape(self.n, -1)
        for s in range(self.first.size - 1):
            f, l = int(self.first[s]), int(self.first[s + 1])
            y = sla.solve_triangular(self.L11[s], x[f:l], lower=True)
            x[f:l] = y
            if self.L21[s].shape[0]:
                below = self.rows[self.rowptr[s]:self.rowptr[s + 1]]
                x[below] -= self.L21[s] @ y
        return x[:, 0] if one else x

    def solve_Lt(self, b, use_LDLt_decomposition=False):
        x = np.array(b, dtype=np.float64)
        one = x.ndim == 1
        x = x.reshape(self.n, -1)
        for s in range(self.first.size - 2, -1, -1):
            f, l = int(self.first[s]), int(self.first[s + 1])
            y = x[f:l]
            if self.L21[s].shape[0]:
                below = self.rows[self.rowptr[s]:self.rowptr[s + 1]]
                y = y - self.L21[s].T @ x[below]
            x[f:l] = sla.solve_triangular(self.L11[s], y, lower=True, trans="T")
        return x[:, 0] if one else x

    def solve_A(self, b):
        b = np.asarray(b.toarray() if sparse.issparse(b) else b, dtype=np.float64)
        return self.apply_Pt(self.solve_Lt(self.solve_L(b[self.perm])))


def factor_with_plan(plan):
    """``impl`` for ``spde_oracle.set_factor``: SupernodalFactor bound to one symbolic plan (falls back to
    the dense factor for matrices of a different size, e.g. the spatial initial-field model)."""
    import spde_oracle as so

    def impl(A, perm=None):
        if A.shape[0] == plan.n:
            return SupernodalFactor(A, plan=plan)
        return so.DenseFactor(A, perm=None)
    return impl
