"""Fixed sparsity pattern of the precision matrices (host side, computed once per mesh).

The device keeps ``Q`` in the slot layouts of ``include/spde_b200.h`` (25 slots per cell for a
spatial model, 43 per node for a space-time model).  This module holds the index arithmetic that
maps those slots to a canonical SciPy CSC matrix -- the form in which the reference exposes
``mod.Q`` (``model.py:135,152``) -- and back.  The reference's own pattern is value dependent
(SciPy's SpGEMM and sparse add drop exact zeros), so parity of indices is
checked after ``eliminate_zeros()`` on both sides.
"""
from __future__ import annotations

import numpy as np
from scipy import sparse


def slot_offsets(T: int, pat: int = 0):
    """(dt, dj, di) of every slot; mirrors ``Geo::slot_offset`` in ``csrc/common.cuh``."""
    if T == 1:
        q = np.arange(25)
        return np.zeros(25, np.int64), q // 5 - 2, q % 5 - 2
    if pat == 1:        # Kronecker pattern of the separable model: 5x5 blocks to t-1, t, t+1
        s = np.arange(75)
        q = s % 25
        return s // 25 - 1, q // 5 - 2, q % 5 - 2
    dt = np.concatenate([np.full(9, -1), np.zeros(25, np.int64), np.full(9, 1)])
    q3, q5 = np.arange(9), np.arange(25)
    dj = np.concatenate([q3 // 3 - 1, q5 // 5 - 2, q3 // 3 - 1])
    di = np.concatenate([q3 % 3 - 1, q5 % 5 - 2, q3 % 3 - 1])
    return dt, dj, di


class Pattern:
    def __init__(self, M: int, N: int, T: int, bc: int, pat: int = 0):
        self.M, self.N, self.T, self.bc, self.pat = M, N, T, bc, pat
        self.Ns = M * N
        self.n = self.Ns * T
        self.nslots = 25 if T == 1 else (75 if pat == 1 else 43)
        node = np.arange(self.n, dtype=np.int64)
        t, k = node // self.Ns, node % self.Ns
        i, j = k % M, k // M
        dt, dj, di = slot_offsets(T, pat)
        ii = i[None, :] + di[:, None]
        jj = j[None, :] + dj[:, None]
        tt = t[None, :] + dt[:, None]
        if bc == 2:
            ii %= M
            jj %= N
            ok = (tt >= 0) & (tt < T)
        else:
            ok = (ii >= 0) & (ii < M) & (jj >= 0) & (jj < N) & (tt >= 0) & (tt < T)
        nbr = tt * self.Ns + jj * M + ii
        self.nbr = np.where(ok, nbr, -1)          # (nslots, n)
        slot, row = np.nonzero(ok)
        col = self.nbr[slot, row]
        order = np.lexsort((row, col))             # CSC: sorted by column, then row
        self.indices = row[order].astype(np.int32)
        self.gather = (slot[order] * self.n + row[order]).astype(np.int64)
        self.indptr = np.concatenate([[0], np.cumsum(np.bincount(col, minlength=self.n))]).astype(np.int32)

    def to_csc(self, flat: np.ndarray) -> sparse.csc_matrix:
        """slot-major values -> canonical CSC on the full geometric pattern (explicit zeros kept)"""
        # indices/indptr are copied: SciPy shares them otherwise and eliminate_zeros() edits in place
        return sparse.csc_matrix((np.asarray(flat)[self.gather], self.indices.copy(), self.indptr.copy()),
                                 shape=(self.n, self.n))

    def from_sparse(self, Q) -> np.ndarray:
        """sparse matrix whose pattern is inside the mesh pattern -> slot-major values"""
        Q = sparse.csc_matrix(Q)
        full = self.to_csc(np.arange(1, self.nslots * self.n + 1, dtype=np.float64))   # position + 1
        pos = full.copy()
        pos.data = np.ones_like(pos.data)
        mask = Q.copy()
        mask.data = np.ones_like(mask.data)
        if (mask - mask.multiply(pos)).nnz:
            raise ValueError("matrix has entries outside the mesh pattern")
        out = np.zeros(self.nslots * self.n)
        Qc = Q.tocoo()
        lookup = full.tocsr()
        idx = np.asarray(lookup[Qc.row, Qc.col]).ravel().astype(np.int64) - 1
        np.add.at(out, idx, Qc.data)
        return out
