"""ctypes binding of ``csrc/libspde_b200.so`` (the C ABI declared in ``include/spde_b200.h``).

The library is the product: there is no CPU fallback.  If it is missing, importing this module
raises; if it is present but no CUDA device is, every numeric entry point fails with a CUDA error
that is turned into :class:`SpdeError` here.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libspde_b200.so")

OK, ERR_NOT_SPD, ERR_OOM, ERR_ARG, ERR_CUDA = 0, 1, 2, 3, 4


class SpdeError(RuntimeError):
    pass


class NotPositiveDefiniteError(SpdeError, ValueError):
    """The reference raises ``CholmodNotPositiveDefiniteError`` here."""


if not os.path.exists(LIB_PATH):
    raise ImportError(
        "%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "(nvcc, sm_100a). spdepy_b200 has no CPU fallback." % LIB_PATH)

lib = ctypes.CDLL(LIB_PATH)

c_int, c_dbl, c_vp, c_i64 = ctypes.c_int, ctypes.c_double, ctypes.c_void_p, ctypes.c_int64

_SIGS = {
    "spde_abi_version": (c_int, []),
    "spde_last_error": (ctypes.c_char_p, []),
    "spde_launch_count": (ctypes.c_longlong, [c_int]),
    "spde_ah_stencil": (c_int, [c_int, c_int, c_int, c_dbl, c_dbl, c_vp, c_int, c_vp, c_vp]),
    "spde_aw_stencil": (c_int, [c_int, c_int, c_int, c_dbl, c_dbl, c_vp, c_vp, c_int, c_int, c_int, c_vp, c_vp]),
    "spde_combine_A": (c_int, [c_int, c_int, c_dbl, c_dbl, c_vp, c_int, c_vp, c_vp, c_vp, c_vp]),
    "spde_atda": (c_int, [c_int, c_int, c_int, c_vp, c_vp, c_int, c_dbl, c_int, c_vp, c_vp]),
    "spde_fill_spacetime": (c_int, [c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_int, c_dbl, c_vp, c_dbl, c_dbl, c_int, c_vp, c_vp]),
    "spde_plan_create": (c_int, [c_int, c_int, c_int, c_int, c_int, ctypes.POINTER(c_vp)]),
    "spde_plan_destroy": (None, [c_vp]),
    "spde_plan_info": (c_i64, [c_vp, c_int]),
    "spde_plan_info_d": (c_dbl, [c_vp, c_int]),
    "spde_plan_perm": (c_int, [c_vp, c_vp]),
    "spde_plan_supernodes": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp]),
    "spde_plan_export": (c_int, [c_vp, c_int, c_int, c_int, c_vp, ctypes.POINTER(c_i64), ctypes.POINTER(c_int)]),
    "spde_plan_profile": (c_int, [c_vp, c_int, c_vp, c_int]),
    "spde_factorize": (c_int, [c_vp, c_int, c_vp, c_vp, c_dbl, c_vp]),
    "spde_factorize_async": (c_int, [c_vp, c_int, c_vp, c_vp, c_dbl, c_vp]),
    "spde_factor_wait": (c_int, [c_vp, c_int, c_vp]),
    "spde_factor_info": (c_int, [c_vp, c_int, ctypes.POINTER(c_int), ctypes.POINTER(c_int)]),
    "spde_logdet": (c_int, [c_vp, c_int, ctypes.POINTER(c_dbl), c_vp]),
    "spde_logdet_dev": (c_int, [c_vp, c_int, c_vp, c_vp]),
    "spde_solve": (c_int, [c_vp, c_int, c_int, c_vp, c_int, c_vp]),
    "spde_selinv": (c_int, [c_vp, c_int, c_vp, c_vp]),
    "spde_selinv_start": (c_int, [c_vp, c_int, c_vp]),
    "spde_selinv_fetch": (c_int, [c_vp, c_int, c_vp, c_vp]),
    "spde_q_apply": (c_int, [c_int, c_int, c_int, c_int, c_vp, c_vp, c_int, c_vp, c_vp]),
    "spde_dot": (c_int, [c_vp, c_vp, c_i64, ctypes.POINTER(c_dbl), c_vp]),
    "spde_wdot": (c_int, [c_vp, c_vp, c_vp, c_i64, c_int, ctypes.POINTER(c_dbl), c_vp]),
    "spde_residual_ss": (c_int, [c_vp, c_vp, c_vp, c_i64, c_int, ctypes.POINTER(c_dbl), c_vp]),
    "spde_dot_dev": (c_int, [c_vp, c_vp, c_i64, c_vp, c_vp]),
    "spde_wdot_dev": (c_int, [c_vp, c_vp, c_vp, c_i64, c_int, c_vp, c_vp]),
    "spde_residual_ss_dev": (c_int, [c_vp, c_vp, c_vp, c_i64, c_int, c_vp, c_vp]),
    "spde_scatter_obs": (c_int, [c_vp, c_vp, c_i64, c_int, c_dbl, c_vp, c_vp]),
    "spde_add_diag": (c_int, [c_vp, c_vp, c_dbl, c_i64, c_vp]),
    "spde_sddmm": (c_int, [c_int, c_int, c_int, c_int, c_vp, c_vp, c_int, c_dbl, c_int, c_vp, c_vp]),
    "spde_assembly_adjoint": (c_int, [c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_int, c_dbl, c_dbl, c_dbl, c_int,
                                      c_vp, c_vp, c_vp, c_vp, c_vp]),
    "spde_gemm_single": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_vp, c_int, c_vp, c_int, c_vp, c_int,
                                 c_int, ctypes.POINTER(ctypes.c_float), c_vp]),
    "spde_stencil_adjoint": (c_int, [c_int, c_int, c_int, c_dbl, c_dbl, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "spde_gemv_t": (c_int, [c_vp, c_vp, c_int, c_int, c_vp, c_vp]),
    "spde_fill_kron": (c_int, [c_int, c_int, c_int, c_int, c_vp, c_dbl, c_dbl, c_dbl, c_vp, c_vp]),
    "spde_kron_reduce": (c_int, [c_int, c_int, c_int, c_int, c_vp, c_dbl, c_dbl, c_dbl, c_vp, c_vp]),
    "spde_potrf_bench": (c_int, [c_int, c_int, c_int, c_int, ctypes.POINTER(ctypes.c_float), c_vp, c_vp]),
    "spde_ooc_create": (c_int, [c_vp, c_i64, c_int, c_int, ctypes.POINTER(c_vp)]),
    "spde_ooc_destroy": (None, [c_vp]),
    "spde_ooc_info": (c_i64, [c_vp, c_int]),
    "spde_ooc_info_d": (c_dbl, [c_vp, c_int]),
    "spde_ooc_run": (c_int, [c_vp, c_vp, c_vp, c_dbl, c_vp, c_int, c_int, c_vp, ctypes.POINTER(c_dbl), c_vp]),
    "spde_ooc_export": (c_int, [c_vp, c_int, c_int, c_int, c_int, c_vp, ctypes.POINTER(c_i64), ctypes.POINTER(c_int)]),
}

EXPORTS = tuple(_SIGS)

for _name, (_res, _args) in _SIGS.items():
    _f = getattr(lib, _name)          # AttributeError here = the .so does not match the header
    _f.restype = _res
    _f.argtypes = _args


def check(rc: int) -> None:
    if rc == OK:
        return
    msg = (lib.spde_last_error() or b"").decode()
    if rc == ERR_NOT_SPD:
        raise NotPositiveDefiniteError(msg)
    if rc == ERR_OOM:
        raise MemoryError(msg)
    if rc == ERR_ARG:
        raise ValueError(msg)
    raise SpdeError("CUDA failure in libspde_b200: %s" % msg)


def ptr(t) -> int:
    """device pointer of a torch tensor (or None)"""
    return None if t is None else t.data_ptr()


def stream_ptr() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream


class PlanHandle:
    """Owner of one ``spde_plan*`` (symbolic analysis + schedules + device workspaces)."""

    def __init__(self, M: int, N: int, T: int, bc: int, pat: int = 0):
        h = c_vp()
        check(lib.spde_plan_create(M, N, T, bc | (pat << 8), 0, ctypes.byref(h)))
        self.h = h
        self.M, self.N, self.T, self.bc, self.pat = M, N, T, bc, pat
        self.n = int(lib.spde_plan_info(h, 0))
        self.nslots = 25 if T == 1 else (75 if pat == 1 else 43)
        self._perm = None

    def info(self, what: int) -> int:
        return int(lib.spde_plan_info(self.h, what))

    def info_d(self, what: int) -> float:
        return float(lib.spde_plan_info_d(self.h, what))

    @property
    def perm(self) -> np.ndarray:
        if self._perm is None:
            p = np.empty(self.n, np.int32)
            check(lib.spde_plan_perm(self.h, p.ctypes.data))
            self._perm = p
        return self._perm

    def supernodes(self):
        ns = self.info(1)
        first = np.empty(ns + 1, np.int32)
        rowptr = np.empty(ns + 1, np.int64)
        rows = np.empty(self.info(9), np.int32)
        parent = np.empty(ns, np.int32)
        check(lib.spde_plan_supernodes(self.h, first.ctypes.data, rowptr.ctypes.data, rows.ctypes.data, parent.ctypes.data))
        return first, rowptr, rows, parent

    def export(self, prog: int, what: int, dtype, k: int = 0) -> np.ndarray:
        """Host copy of a schedule / layout array (``spde_plan_export``)."""
        cnt, es = c_i64(), c_int()
        check(lib.spde_plan_export(self.h, prog, k, what, None, ctypes.byref(cnt), ctypes.byref(es)))
        dtype = np.dtype(dtype)
        if cnt.value and dtype.itemsize != es.value:
            raise SpdeError("export dtype size %d != %d" % (dtype.itemsize, es.value))
        out = np.empty(cnt.value, dtype)
        if cnt.value:
            check(lib.spde_plan_export(self.h, prog, k, what, out.ctypes.data, None, None))
        return out

    def profile(self, enable: bool, reset: bool = True):
        """Toggle per-launch timing; returns (ms, counts) arrays [kind, variant] accumulated so far."""
        out = np.zeros(2 * 8 * 16)
        check(lib.spde_plan_profile(self.h, int(enable), out.ctypes.data, int(reset)))
        return out[:128].reshape(8, 16), out[128:].reshape(8, 16)

    def stats(self) -> dict:
        return {"n": self.n, "nsuper": self.info(1), "nnzL": self.info(2), "flops": self.info_d(3),
                "factor_bytes": self.info(4), "arena_bytes": self.info(5), "levels": self.info(6),
                "max_front": self.info(7), "max_cols": self.info(13), "launches": self.info(8),
                "gemm_tasks": self.info(11), "tiles": self.info(12), "zarena_bytes": self.info(10),
                "sched_flops": self.info_d(14), "dinv_bytes": self.info(15), "ybuf_bytes": self.info(16)}

    def __del__(self):
        try:
            if self.h:
                lib.spde_plan_destroy(self.h)
                self.h = None
        except Exception:
            pass


class OocHandle:
    """Owner of one ``spde_ooc*``: the streamed (depth-first, statically planned) evaluator of a plan, for meshes
    whose factor does not fit in HBM (``include/spde_b200.h``, "streamed evaluation")."""

    def __init__(self, plan: PlanHandle, top_bytes: int, backward: bool = True, build: bool = True):
        h = c_vp()
        check(lib.spde_ooc_create(plan.h, int(top_bytes), int(backward), int(build), ctypes.byref(h)))
        self.h = h
        self.plan = plan            # keeps the plan alive
        self.top_bytes, self.backward, self.built = int(top_bytes), bool(backward), bool(build)

    def info(self, what: int) -> int:
        return int(lib.spde_ooc_info(self.h, what))

    def info_d(self, what: int) -> float:
        return float(lib.spde_ooc_info_d(self.h, what))

    def stats(self) -> dict:
        return {"segments": self.info(0), "top_segments": self.info(1), "pool_bytes": self.info(2),
                "peak_forward_bytes": self.info(3), "peak_backward_bytes": self.info(4), "host_bytes": self.info(5),
                "max_working_set_bytes": self.info(7), "factor_launches": self.info(8), "selinv_launches": self.info(9),
                "recompute_flops": self.info_d(0), "top_bytes": self.top_bytes}

    def export(self, seg: int, prog: int, what: int, dtype, k: int = 0) -> np.ndarray:
        cnt, es = c_i64(), c_int()
        check(lib.spde_ooc_export(self.h, seg, prog, k, what, None, ctypes.byref(cnt), ctypes.byref(es)))
        dtype = np.dtype(dtype)
        if cnt.value and dtype.itemsize != es.value:
            raise SpdeError("export dtype size %d != %d" % (dtype.itemsize, es.value))
        out = np.empty(cnt.value, dtype)
        if cnt.value:
            check(lib.spde_ooc_export(self.h, seg, prog, k, what, out.ctypes.data, None, None))
        return out

    def run(self, Q_ptr, cnt_ptr, tau: float, X_ptr, k: int, mode: int, Zq_ptr, stream) -> float:
        ld = c_dbl()
        check(lib.spde_ooc_run(self.h, Q_ptr, cnt_ptr, float(tau), X_ptr, int(k), int(mode), Zq_ptr, ctypes.byref(ld), stream))
        return ld.value

    def __del__(self):
        try:
            if self.h:
                lib.spde_ooc_destroy(self.h)
                self.h = None
        except Exception:
            pass
