"""Oracle-side symbolic analysis: fill-reducing permutation + supernodal structure.  TEST INFRASTRUCTURE ONLY.

The reference's ordering and symbolic factorisation happen inside SuiteSparse CHOLMOD (scikit-sparse 0.4.12,
``poetry.lock:582-583``; every ``cholesky(Q)`` call analyses again, ``advection_diffusion2D.py:117,193``), which is
absent from the image.  This module gives the oracle's CPU Cholesky (``oracle/cpu_cholesky.py``) a symbolic phase of
its own, so that ``bench.py --impl reference`` never touches ``spdepy_b200``:

* :func:`nd_perm` -- geometric nested dissection of the ``M x N x T`` mesh (George 1973): separators two cells thick
  in x / y and one slice thick in t, which is what the 5x5 in-slice / 3x3 slice-to-slice coupling of Q needs
  (SURVEY.md App. D); leaves of <= ``leaf`` nodes are ordered naturally.
* :class:`OracleSymbolic` -- elimination tree, postorder, column counts, relaxed supernodes and their row structures
  from ``oracle/symbolic_oracle.c`` (restating CHOLMOD's published symbolic phase on the general CSC pattern).

It exposes the three things ``cpu_cholesky.SupernodalFactor`` reads from a symbolic plan: ``n``, ``perm`` (new -> old)
and ``supernodes()``.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np
from scipy import sparse

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle_symbolic.so")
        src = os.path.join(_HERE, "symbolic_oracle.c")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", _HERE, "liboracle_symbolic.so"], stdout=subprocess.DEVNULL)
        L = ctypes.CDLL(path)
        L.osym_analyse.restype = ctypes.c_void_p
        L.osym_analyse.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        L.osym_free.argtypes = [ctypes.c_void_p]
        L.osym_nsuper.argtypes = [ctypes.c_void_p]
        L.osym_nrows.argtypes = [ctypes.c_void_p]
        L.osym_nrows.restype = ctypes.c_int64
        L.osym_nnzL.argtypes = [ctypes.c_void_p]
        L.osym_nnzL.restype = ctypes.c_double
        L.osym_flops.argtypes = [ctypes.c_void_p]
        L.osym_flops.restype = ctypes.c_double
        L.osym_export.argtypes = [ctypes.c_void_p] * 6
        L.osym_extend_add.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                                      ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]
        _LIB = L
    return _LIB


def nd_perm(M: int, N: int, T: int = 1, bc: int = 3, leaf: int | None = None) -> np.ndarray:
    """Nested-dissection ordering (new -> old) of the nodes ``k = t*M*N + y*M + x``."""
    Ns = M * N
    leaf = (32 if T == 1 else 64) if leaf is None else leaf
    out = []

    def emit(x0, x1, y0, y1, t0, t1):
        if x1 > x0 and y1 > y0 and t1 > t0:
            t, y, x = np.meshgrid(np.arange(t0, t1), np.arange(y0, y1), np.arange(x0, x1), indexing="ij")
            out.append((t * Ns + y * M + x).reshape(-1))

    def nd(x0, x1, y0, y1, t0, t1):
        lens = (x1 - x0, y1 - y0, t1 - t0)
        vol = lens[0] * lens[1] * lens[2]
        if vol <= 0:
            return
        best = None
        if vol > leaf:
            for d, thick in enumerate((2, 2, 1)):
                if lens[d] < thick + 2:
                    continue
                cost = thick * (vol // lens[d])
                if best is None or cost < best[1] or (cost == best[1] and lens[d] > lens[best[0]]):
                    best = (d, cost)
        if best is None:
            emit(x0, x1, y0, y1, t0, t1)
            return
        d = best[0]
        if d == 0:
            mid = x0 + (lens[0] - 2) // 2
            nd(x0, mid, y0, y1, t0, t1); nd(mid + 2, x1, y0, y1, t0, t1); emit(mid, mid + 2, y0, y1, t0, t1)
        elif d == 1:
            mid = y0 + (lens[1] - 2) // 2
            nd(x0, x1, y0, mid, t0, t1); nd(x0, x1, mid + 2, y1, t0, t1); emit(x0, x1, mid, mid + 2, t0, t1)
        else:
            mid = t0 + (lens[2] - 1) // 2
            nd(x0, x1, y0, y1, t0, mid); nd(x0, x1, y0, y1, mid + 1, t1); emit(x0, x1, y0, y1, mid, mid + 1)

    if bc == 2:     # periodic in x and y: the wrap couples the two ends, so the strips x < 2 and y < 2 go last
        nd(2, M, 2, N, 0, T)
        emit(0, 2, 2, N, 0, T)
        emit(0, M, 0, 2, 0, T)
    else:
        nd(0, M, 0, N, 0, T)
    perm = np.concatenate(out).astype(np.int64)
    assert perm.size == Ns * T
    return perm


def extend_add(U, rel, nc, F11, F21, F22):
    """``F[rel, rel] += U`` on the lower triangle (``U``: child update matrix, column-major; ``rel``: ascending positions
    of its rows in the parent front whose blocks are ``F11`` (nc x nc), ``F21`` (nr x nc), ``F22`` (nr x nr), column-major)."""
    rel = np.ascontiguousarray(rel, dtype=np.int32)
    _lib().osym_extend_add(U.ctypes.data, U.shape[0], U.shape[0], rel.ctypes.data, nc, F11.ctypes.data,
                           F21.ctypes.data if F21 is not None else None, 0 if F22 is None else F22.shape[0],
                           F22.ctypes.data if F22 is not None else None)


class OracleSymbolic:
    """Symbolic analysis of ``P A P^T`` for a given ordering ``perm`` (new -> old).  The final permutation is ``perm``
    composed with a postorder of the elimination tree."""

    def __init__(self, A, perm, relax: bool = True):
        A = sparse.csc_matrix(A)
        n = A.shape[0]
        self.n = n
        perm = np.asarray(perm, dtype=np.int64)
        ip = np.empty(n, np.int64)
        ip[perm] = np.arange(n)
        C = A.tocoo()
        # symmetric pattern of the permuted matrix (both triangles)
        r, c = ip[C.row], ip[C.col]
        B = sparse.csc_matrix((np.ones(2 * r.size, dtype=np.int8), (np.concatenate([r, c]), np.concatenate([c, r]))), shape=(n, n))
        Ap = np.ascontiguousarray(B.indptr, dtype=np.int64)
        Ai = np.ascontiguousarray(B.indices, dtype=np.int32)
        L = _lib()
        h = L.osym_analyse(n, Ap.ctypes.data, Ai.ctypes.data, int(relax))
        try:
            ns, nr = L.osym_nsuper(h), L.osym_nrows(h)
            post = np.empty(n, np.int32)
            self.first = np.empty(ns + 1, np.int32)
            self.rowptr = np.empty(ns + 1, np.int64)
            self.rows = np.empty(nr, np.int32)
            self.sparent = np.empty(ns, np.int32)
            L.osym_export(h, post.ctypes.data, self.first.ctypes.data, self.rowptr.ctypes.data, self.rows.ctypes.data,
                          self.sparent.ctypes.data)
            self.nnzL, self.flops = L.osym_nnzL(h), L.osym_flops(h)
        finally:
            L.osym_free(h)
        self.perm = perm[post]
        self.nsuper = ns

    def supernodes(self):
        return self.first, self.rowptr, self.rows, self.sparent

    def stats(self):
        return {"n": self.n, "nsuper": self.nsuper, "nnzL": self.nnzL, "flops": self.flops}
