"""Run the UNMODIFIED reference (berild/spdepy) in this container.  TEST INFRASTRUCTURE ONLY.

The reference cannot travel to the GPU box (``/root/reference`` is absent there), so this module
is used in exactly two places: ``oracle/make_golden.py`` (writes ``tests/golden/*.npz``) and the
``not gpu`` tests that cross-check ``oracle/spde_oracle.py`` when the reference tree is mounted.
Nothing under ``spdepy_b200/`` imports it.

What it does:
  * makes a scratch copy of ``/root/reference/src/spdepy`` under ``/tmp`` (the reference writes
    its ``.so`` files next to its sources, ``advection_diffusion2D.py:262-304``) and drops the
    libraries built by ``oracle/Makefile`` (``oracle/_ref/lib_*.so``) into ``ccode/`` so that the
    reference's broken Linux link line (``advection_diffusion2D.py:269-270``) is never reached;
  * shims three imports: ``importlib.metadata.version("spdepy")`` (``src/spdepy/__init__.py:2-3``),
    ``netCDF4`` (``datasets.py:1``) and ``sksparse.cholmod`` (every ``spdes/*.py`` line 5).

The ``sksparse.cholmod`` stand-in factorises ``P Q P^T`` densely with LAPACK; ``P`` is either the
identity or a permutation injected with :func:`set_permutation` so that ``P^T L^-T z`` samples are
comparable with a build that uses the same ``P`` (SURVEY.md finding 3).  scikit-sparse 0.4.12 /
SuiteSparse CHOLMOD themselves are absent from the image: parity at that boundary is "unpinned"
by the reference's own tests and pinned here only through this stand-in.
"""
from __future__ import annotations

import importlib
import importlib.metadata
import os
import shutil
import sys
import types

import numpy as np
from scipy import sparse
from scipy import linalg as sla

REF_ROOT = "/root/reference"
_HERE = os.path.dirname(os.path.abspath(__file__))
_SCRATCH = "/tmp/spdepy_ref_scratch_%d" % os.getuid()

_perm_for_n: dict[int, np.ndarray] = {}


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "src", "spdepy"))


def set_permutation(perm: np.ndarray | None, n: int | None = None) -> None:
    """Use ``perm`` (new -> old) for every later factorisation of an ``n x n`` matrix."""
    if perm is None:
        if n is None:
            _perm_for_n.clear()
        else:
            _perm_for_n.pop(n, None)
        return
    perm = np.asarray(perm, dtype=np.int64)
    _perm_for_n[perm.size if n is None else n] = perm


class DenseFactor:
    """Stand-in for ``sksparse.cholmod.Factor`` (methods used at ``advection_diffusion2D.py:
    194-202`` and ``model.py:80,126``).  ``L L^T = P A P^T`` with ``(P x)[i] = x[perm[i]]``."""

    def __init__(self, A):
        A = sparse.csc_matrix(A)
        n = A.shape[0]
        self.n = n
        self.perm = _perm_for_n.get(n, np.arange(n, dtype=np.int64))
        # CHOLMOD reads the lower triangle only
        Al = sparse.tril(A).toarray()
        Ad = Al + np.tril(Al, -1).T
        Ap = Ad[np.ix_(self.perm, self.perm)]
        self.L = np.linalg.cholesky(Ap)

    def P(self):
        return self.perm.copy()

    def logdet(self):
        return 2.0 * np.log(np.diag(self.L)).sum()

    def apply_P(self, x):
        return np.asarray(x)[self.perm]

    def apply_Pt(self, x):
        x = np.asarray(x)
        out = np.empty_like(x)
        out[self.perm] = x
        return out

    def solve_L(self, b, use_LDLt_decomposition=False):
        return sla.solve_triangular(self.L, np.asarray(b, dtype=np.float64), lower=True)

    def solve_Lt(self, b, use_LDLt_decomposition=False):
        return sla.solve_triangular(self.L, np.asarray(b, dtype=np.float64), lower=True, trans="T")

    def solve_A(self, b):
        b = np.asarray(b)
        if sparse.issparse(b):
            b = b.toarray()
        b = np.asarray(b, dtype=np.float64)
        y = sla.solve_triangular(self.L, b[self.perm], lower=True)
        x = sla.solve_triangular(self.L, y, lower=True, trans="T")
        return self.apply_Pt(x)

    __call__ = solve_A


_sparse_impl = None


def set_sparse_factor(impl) -> None:
    """``impl(A) -> Factor`` used for matrices too large for the dense stand-in (n > 12000): the BASELINE-config
    fixtures of ``oracle/make_golden_baseline.py`` pass ``cpu_cholesky.SupernodalFactor`` bound to an
    ``OracleSymbolic`` plan.  ``None`` restores the dense-only behaviour."""
    global _sparse_impl
    _sparse_impl = impl


def _cholesky(A, **kwargs):
    if sparse.issparse(A) and A.shape[0] > 12000:
        if _sparse_impl is None:
            raise RuntimeError("dense CHOLMOD stand-in limited to n <= 12000 (see set_sparse_factor)")
        return _sparse_impl(A)
    return DenseFactor(A)


def _install_shims() -> None:
    if "sksparse.cholmod" not in sys.modules:
        pkg = types.ModuleType("sksparse")
        mod = types.ModuleType("sksparse.cholmod")
        mod.cholesky = _cholesky
        mod.Factor = DenseFactor
        mod.CholmodNotPositiveDefiniteError = np.linalg.LinAlgError
        pkg.cholmod = mod
        sys.modules["sksparse"] = pkg
        sys.modules["sksparse.cholmod"] = mod
    if "netCDF4" not in sys.modules:
        nc = types.ModuleType("netCDF4")
        nc.Dataset = object
        sys.modules["netCDF4"] = nc
    if not getattr(importlib.metadata, "_spdepy_shim", False):
        orig = importlib.metadata.version

        def version(name):
            if name == "spdepy":
                return "0.1.0"
            return orig(name)

        importlib.metadata.version = version
        importlib.metadata._spdepy_shim = True


def load_reference():
    """Import and return the reference package ``spdepy`` (unmodified sources, scratch copy)."""
    if "spdepy" in sys.modules and getattr(sys.modules["spdepy"], "_is_reference", False):
        return sys.modules["spdepy"]
    if not available():
        raise RuntimeError("reference tree not mounted at %s" % REF_ROOT)
    src = os.path.join(REF_ROOT, "src", "spdepy")
    dst = os.path.join(_SCRATCH, "spdepy")
    if not os.path.isdir(dst):
        os.makedirs(_SCRATCH, exist_ok=True)
        shutil.copytree(src, dst)
    refdir = os.path.join(_HERE, "_ref")
    if not os.path.isdir(refdir) or not os.listdir(refdir):
        os.system("make -C %s ref > /dev/null" % _HERE)
    for f in os.listdir(refdir):
        if f.endswith(".so"):
            tgt = os.path.join(dst, "spdes", "ccode", f)
            if not os.path.exists(tgt):
                shutil.copy(os.path.join(refdir, f), tgt)
    _install_shims()
    if _SCRATCH not in sys.path:
        sys.path.insert(0, _SCRATCH)
    sp = importlib.import_module("spdepy")
    sp._is_reference = True
    return sp


def seeded_probes(n: int, nh1: int, seed: int) -> np.ndarray:
    """The reference's probe draw (``advection_diffusion2D.py:200``) with the global legacy RNG
    seeded immediately before (SURVEY.md section 8c "Seeding")."""
    np.random.seed(seed)
    return (2 * np.random.randint(1, 3, n * nh1) - 3).reshape(n, nh1)


def loglike_seeded(mod, par, nh1=100, grad=True, seed=4):
    np.random.seed(seed)
    return mod.logLike(np.asarray(par, dtype="float64"), nh1=nh1, grad=grad)
