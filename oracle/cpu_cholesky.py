"""CPU supernodal multifrontal Cholesky -- the oracle's stand-in for CHOLMOD.  TEST INFRASTRUCTURE ONLY.

The reference factorises through scikit-sparse 0.4.12 -> SuiteSparse CHOLMOD (``poetry.lock:582-583``;
call sites ``advection_diffusion2D.py:117,193``, ``model.py:79,125``), a system library that is absent
from ``/root/reference`` and from this image.  This module restates CHOLMOD's published supernodal
algorithm (Chen, Davis, Hager, Rajamanickam, ACM TOMS 35(3), 2008: supernodal ``L L^T`` with dense
``potrf / trsm / syrk`` on the supernodes) in multifrontal form with NumPy / LAPACK dense kernels.
The fill-reducing permutation and the supernode partition come from a symbolic plan object (``n``,
``perm``, ``supernodes()``): the tests pass the build's own analysis (``spde_plan_perm`` /
``spde_plan_supernodes``, host code), because ``P^T L^-T z`` samples are only comparable under the same
``P`` (SURVEY.md finding 3); ``bench.py --impl reference`` passes ``oracle/symbolic_oracle.OracleSymbolic``
(nested dissection + CHOLMOD's published symbolic phase restated in C), so that arm never loads the product.

It is used (a) as the independent numeric check of the CUDA factorisation at sizes where a dense
factor is too big, and (b) as the CPU baseline that ``bench.py`` times on the GPU box's host cores.
PARITY STATUS: no reference test pins numbers at the CHOLMOD boundary; this stand-in
is itself checked against dense LAPACK in tests/test_cpu_cholesky.py.
"""
from __future__ import annotations

import numpy as np
from scipy import linalg as sla
from scipy import sparse

import symbolic_oracle as syo


class SupernodalFactor:
    """``L L^T = P A P^T``; same method surface as ``sksparse.cholmod.Factor`` as used by the
    reference: ``logdet, solve_A, solve_Lt, solve_L, apply_P, apply_Pt, P``."""

    def __init__(self, A, perm=None, plan=None):
        if plan is None:
            raise ValueError("SupernodalFactor needs a symbolic plan: the build's PlanHandle (tests: samples under the "
                             "same P) or oracle/symbolic_oracle.OracleSymbolic (bench.py's reference arm)")
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
        dpotrf, dtrsm, dsyrk = sla.lapack.dpotrf, sla.blas.dtrsm, sla.blas.dsyrk
        for s in range(ns):
            f, l = int(first[s]), int(first[s + 1])
            nc = l - f
            below = self.rows[rowptr[s]:rowptr[s + 1]]
            nr = below.size
            # the frontal matrix in three column-major blocks; only lower triangles are ever read
            F11 = np.zeros((nc, nc), order="F")
            F21 = np.zeros((nr, nc), order="F")
            F22 = np.zeros((nr, nr), order="F") if nr else None
            # original entries of the pivot columns
            lo_p, hi_p = indptr[f], indptr[l]
            ri = indices[lo_p:hi_p]
            cj = np.repeat(np.arange(nc), np.diff(indptr[f:l + 1]))
            piv = ri < l
            F11[ri[piv] - f, cj[piv]] = data[lo_p:hi_p][piv]
            if nr:
                F21[np.searchsorted(below, ri[~piv]), cj[~piv]] = data[lo_p:hi_p][~piv]
            # extend-add of the children's update matrices (their row lists are ascending, so the entries that land
            # in the pivot block come first)
            for c_ in kids[s]:
                cb = self.rows[rowptr[c_]:rowptr[c_ + 1]]
                rel = np.where(cb < l, cb - f, nc + np.searchsorted(below, cb))
                syo.extend_add(upd[c_], rel, nc, F11, F21 if nr else None, F22)
                upd[c_] = None
            L11, info = dpotrf(F11, lower=1, overwrite_a=1, clean=1)
            if info != 0:
                raise np.linalg.LinAlgError("matrix is not positive definite (pivot %d of the permuted matrix)" % (f + info - 1))
            self.L11[s] = L11
            if nr:
                L21 = dtrsm(1.0, L11, F21, side=1, lower=1, trans_a=1, overwrite_b=1)
                self.L21[s] = L21
                upd[s] = dsyrk(-1.0, L21, beta=1.0, c=F22, lower=1, overwrite_c=1)
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

    # The right-hand sides are kept node-major ((n, k), C order): the rows of one supernode are then a contiguous
    # (nc x k) block whose transpose is a column-major (k x nc) matrix, so the triangular solves and the updates run in
    # place through BLAS dtrsm / dgemm "from the right" (Y^T L^T = B^T) without copies or finiteness scans.  Every dense
    # call goes through scipy.linalg.blas -- mixing it with numpy's matmul would alternate between two OpenBLAS pools
    # whose spinning workers fight for the cores.
    def solve_L(self, b, use_LDLt_decomposition=False):
        x = np.array(b, dtype=np.float64, order="C")
        one = x.ndim == 1
        x = x.reshape(self.n, -1)
        dtrsm, dgemm = sla.blas.dtrsm, sla.blas.dgemm
        for s in range(self.first.size - 1):
            f, l = int(self.first[s]), int(self.first[s + 1])
            xt = x[f:l].T                                   # (k x nc), column-major view
            dtrsm(1.0, self.L11[s], xt, side=1, lower=1, trans_a=1, overwrite_b=1)
            if self.L21[s].shape[0]:
                below = self.rows[self.rowptr[s]:self.rowptr[s + 1]]
                x[below] -= dgemm(1.0, xt, self.L21[s], trans_b=1).T      # (L21 y)^T = y^T L21^T
        return x[:, 0] if one else x

    def solve_Lt(self, b, use_LDLt_decomposition=False):
        x = np.array(b, dtype=np.float64, order="C")
        one = x.ndim == 1
        x = x.reshape(self.n, -1)
        dtrsm, dgemm = sla.blas.dtrsm, sla.blas.dgemm
        for s in range(self.first.size - 2, -1, -1):
            f, l = int(self.first[s]), int(self.first[s + 1])
            xt = x[f:l].T
            if self.L21[s].shape[0]:
                below = self.rows[self.rowptr[s]:self.rowptr[s + 1]]
                dgemm(-1.0, x[below].T, self.L21[s], beta=1.0, c=xt, overwrite_c=1)       # x_s^T -= x_below^T L21
            dtrsm(1.0, self.L11[s], xt, side=1, lower=1, trans_a=0, overwrite_b=1)
        return x[:, 0] if one else x

    def solve_A(self, b):
        b = np.asarray(b.toarray() if sparse.issparse(b) else b, dtype=np.float64)
        return self.apply_Pt(self.solve_Lt(self.solve_L(b[self.perm])))

    __call__ = solve_A


def factor_with_plan(plan):
    """``impl`` for ``spde_oracle.set_factor``: SupernodalFactor bound to one symbolic plan (falls back to
    the dense factor for matrices of a different size, e.g. the spatial initial-field model)."""
    import spde_oracle as so

    def impl(A, perm=None):
        if A.shape[0] == plan.n:
            return SupernodalFactor(A, plan=plan)
        return so.DenseFactor(A, perm=None)
    return impl
