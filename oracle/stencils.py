"""ctypes access to the stencil oracle and to the compiled reference stencils.

TEST INFRASTRUCTURE ONLY: imported by tests/, bench.py's cpu_baseline leg and
__graft_entry__.smoke() -- never by spdepy_b200/.

``oracle_*``  call ``oracle/liboracle_stencil.so`` (C restatement, oracle/stencil_oracle.c).
``ref_*``     call ``oracle/_ref/lib_<name>_b<bc>.so`` = the reference's own C++ compiled from
              ``/root/reference/src/spdepy/spdes/ccode`` by ``oracle/Makefile`` (C-ABI at e.g.
              ``AcH_2D_b1.cpp:170-185``, ``Aw_2D_b1.cpp:136-151``).
Both return the raw cell-major COO triplets ``(row, col, val)``; ``to_csc`` applies the filter
and conversion of the reference's Python wrappers (``advection_diffusion2D.py:226-260``,
``var_advection_var_diffusion2D.py:239-276``).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np
from scipy import sparse

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None
_ref_libs: dict[str, ctypes.CDLL] = {}

_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


def build(quiet: bool = True) -> None:
    subprocess.run(["make", "-C", _HERE, "all"], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def _oracle():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "liboracle_stencil.so")
        if not os.path.exists(path):
            build()
        _lib = ctypes.CDLL(path)
        _lib.orc_ah_const.argtypes = [ctypes.c_int, ctypes.c_int, _dp, ctypes.c_double, ctypes.c_double,
                                      ctypes.c_int, _ip, _ip, _dp]
        _lib.orc_ah_face.argtypes = [ctypes.c_int, ctypes.c_int, _dp, ctypes.c_double, ctypes.c_double,
                                     ctypes.c_int, _ip, _ip, _dp, _dp]
        _lib.orc_aw_const.argtypes = [ctypes.c_int, ctypes.c_int, _dp, ctypes.c_double, ctypes.c_double,
                                      ctypes.c_int, ctypes.c_int, _ip, _ip, _dp]
        _lib.orc_aw_face.argtypes = [ctypes.c_int, ctypes.c_int, _dp, _dp, ctypes.c_double, ctypes.c_double,
                                     ctypes.c_int, ctypes.c_int, _ip, _ip, _dp, _dp]
    return _lib


def _alloc(ns, w):
    return (np.empty(ns * w, np.int32), np.empty(ns * w, np.int32), np.empty(ns * w, np.float64))


def oracle_ah_const(M, N, H, hx, hy, bc):
    row, col, val = _alloc(M * N, 9)
    rc = _oracle().orc_ah_const(M, N, np.ascontiguousarray(H, np.float64).reshape(4), hx, hy, bc, row, col, val)
    if rc:
        raise ValueError("constant-H stencil with bc=2 is undefined in the reference (AcH_2D_b2.cpp:105)")
    return row, col, val


def oracle_ah_face(M, N, H, hx, hy, bc):
    row, col, val = _alloc(M * N, 9)
    H = np.ascontiguousarray(H, np.float64).reshape(M * N * 16)
    _oracle().orc_ah_face(M, N, H, hx, hy, bc, row, col, val, np.empty(M * N * 16))
    return row, col, val


def oracle_aw_const(M, N, G, hx, hy, diff, bc):
    row, col, val = _alloc(M * N, 5)
    _oracle().orc_aw_const(M, N, np.ascontiguousarray(G, np.float64).reshape(2), hx, hy, diff, bc, row, col, val)
    return row, col, val


def oracle_aw_face(M, N, G, dG, hx, hy, diff, bc):
    row, col, val = _alloc(M * N, 5)
    G = np.ascontiguousarray(G, np.float64).reshape(M * N * 4)
    dG = np.zeros(M * N * 4) if dG is None else np.ascontiguousarray(dG, np.float64).reshape(M * N * 4)
    _oracle().orc_aw_face(M, N, G, dG, hx, hy, diff, bc, row, col, val, np.empty(M * N * 8))
    return row, col, val


# ---------------------------------------------------------------------------------------------
# the compiled reference (oracle/_ref)

def ref_available() -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", "lib_AcH_2D_b1.so"))


def _ref(name):
    if name not in _ref_libs:
        _ref_libs[name] = ctypes.CDLL(os.path.join(_HERE, "_ref", "lib_%s.so" % name))
    return _ref_libs[name]


def _ref_fetch(lib, prefix, obj, ns, w):
    out = []
    for fn, dt in (("Row", ctypes.c_int), ("Col", ctypes.c_int), ("Val", ctypes.c_double)):
        f = getattr(lib, "%s_%s" % (prefix, fn))
        f.argtypes = [ctypes.c_void_p]
        f.restype = ctypes.POINTER(dt)
        out.append(np.ctypeslib.as_array(f(obj), shape=(ns * w,)).copy())
    d = getattr(lib, "%s_delete" % prefix)
    d.argtypes = [ctypes.c_void_p]
    d.restype = None
    d(obj)
    return tuple(out)


def ref_ah_const(M, N, H, hx, hy, bc):
    lib = _ref("AcH_2D_b%d" % bc)
    lib.AH_new.argtypes = [ctypes.c_int, ctypes.c_int, _dp, ctypes.c_double, ctypes.c_double]
    lib.AH_new.restype = ctypes.c_void_p
    obj = lib.AH_new(M, N, np.array(H, np.float64).reshape(4), hx, hy)
    return _ref_fetch(lib, "AH", obj, M * N, 9)


def ref_ah_face(M, N, H, hx, hy, bc):
    lib = _ref("AH_2D_b%d" % bc)
    lib.AH_new.argtypes = [ctypes.c_int, ctypes.c_int, _dp, ctypes.c_double, ctypes.c_double]
    lib.AH_new.restype = ctypes.c_void_p
    obj = lib.AH_new(M, N, np.array(H, np.float64).reshape(M * N * 16), hx, hy)   # private copy (b1 mutates)
    return _ref_fetch(lib, "AH", obj, M * N, 9)


def ref_aw_const(M, N, G, hx, hy, diff, bc):
    lib = _ref("Acw_2D_b%d" % bc)
    lib.Aw_new.argtypes = [ctypes.c_int, ctypes.c_int, _dp, ctypes.c_double, ctypes.c_double, ctypes.c_int]
    lib.Aw_new.restype = ctypes.c_void_p
    obj = lib.Aw_new(M, N, np.array(G, np.float64).reshape(2), hx, hy, diff)
    return _ref_fetch(lib, "Aw", obj, M * N, 5)


def ref_aw_face(M, N, G, dG, hx, hy, diff, bc):
    lib = _ref("Aw_2D_b%d" % bc)
    lib.Aw_new.argtypes = [ctypes.c_int, ctypes.c_int, _dp, ctypes.c_double, ctypes.c_double, ctypes.c_int, _dp]
    lib.Aw_new.restype = ctypes.c_void_p
    G = np.array(G, np.float64).reshape(M * N * 4)
    dG = np.zeros(M * N * 4) if dG is None else np.array(dG, np.float64).reshape(M * N * 4)
    obj = lib.Aw_new(M, N, G, hx, hy, diff, dG)
    return _ref_fetch(lib, "Aw", obj, M * N, 5)


def to_csc(triplet, ns, nan_to_zero=False):
    """``row != M*N`` filter + ``csc_matrix((val,(row,col)))`` (duplicates summed), as in
    ``advection_diffusion2D.py:254-258``; ``nan_to_zero`` as ``var_advection_var_diffusion2D.py:255``."""
    row, col, val = triplet
    keep = row != ns
    row, col, val = row[keep], col[keep], val[keep]
    if nan_to_zero:
        val = val.copy()
        val[np.isnan(val)] = 0.0
    return sparse.csc_matrix((val, (row, col)), shape=(ns, ns))
