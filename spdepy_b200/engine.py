"""Device engine: thin, stateful Python over the C ABI (``include/spde_b200.h``).

One :class:`Engine` per (mesh shape, boundary condition).  It owns the symbolic plan (created
lazily, once per mesh -- the reference re-analyses on every ``cholesky`` call,
``advection_diffusion2D.py:117,193``), the fixed sparsity pattern, and convenience wrappers that
take/return ``torch.cuda.DoubleTensor`` buffers.  PyTorch is used for device memory and streams
only; every numeric step is a kernel of ``libspde_b200.so``.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np
import torch

from . import _lib
from ._lib import check, lib, ptr
from .pattern import Pattern

F64 = torch.float64


def _dev():
    if not torch.cuda.is_available():
        raise _lib.SpdeError("spdepy_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _stream():
    return torch.cuda.current_stream().cuda_stream


# bytes moved across PCIe and library calls issued, for bench.py's h2d/d2h/launch accounting
COUNTERS = {"h2d": 0, "d2h": 0, "calls": 0}


class nvtx_range:
    """NVTX range around a phase of the hot path (assembly, factorisation, solves, Takahashi, gradient contraction) when
    ``SPDE_NVTX=1``: the ranges show up in any CUDA timeline tool; a no-op otherwise."""
    enabled = os.environ.get("SPDE_NVTX", "0") not in ("0", "")

    def __init__(self, name: str):
        self.name = name

    def __enter__(self):
        if nvtx_range.enabled:
            torch.cuda.nvtx.range_push(self.name)
        return self

    def __exit__(self, *a):
        if nvtx_range.enabled:
            torch.cuda.nvtx.range_pop()


def to_dev(a, dtype=F64) -> torch.Tensor:
    if isinstance(a, torch.Tensor):
        if not a.is_cuda:
            COUNTERS["h2d"] += a.numel() * a.element_size()
        return a.to(device=_dev(), dtype=dtype, non_blocking=True).contiguous()
    a = np.ascontiguousarray(a)
    COUNTERS["h2d"] += a.nbytes
    return torch.as_tensor(a, dtype=dtype, device=_dev())


def to_host(t: torch.Tensor) -> np.ndarray:
    COUNTERS["d2h"] += t.numel() * t.element_size()
    return t.cpu().numpy()


class Lazy:
    """A scalar that lives in a :class:`ScalarPool` (or a linear expression of such scalars and host numbers): the
    arithmetic of ``logLike`` is written as in the reference, but nothing is read from the device until the first
    ``float()`` -- which fetches the whole pool in ONE device-to-host copy."""
    __slots__ = ("pool", "fn")
    __array_ufunc__ = None          # NumPy scalars defer to the reflected operators below

    def __init__(self, pool, fn):
        self.pool, self.fn = pool, fn

    def __float__(self):
        return float(self.fn(self.pool.fetch()))

    @staticmethod
    def _f(o):
        return o.fn if isinstance(o, Lazy) else (lambda h, v=float(o): v)

    def __add__(self, o):
        a, b = self.fn, Lazy._f(o)
        return Lazy(self.pool, lambda h: a(h) + b(h))
    __radd__ = __add__

    def __sub__(self, o):
        a, b = self.fn, Lazy._f(o)
        return Lazy(self.pool, lambda h: a(h) - b(h))

    def __rsub__(self, o):
        a, b = self.fn, Lazy._f(o)
        return Lazy(self.pool, lambda h: b(h) - a(h))

    def __mul__(self, o):
        a, b = self.fn, Lazy._f(o)
        return Lazy(self.pool, lambda h: a(h) * b(h))
    __rmul__ = __mul__

    def __truediv__(self, o):
        a, b = self.fn, Lazy._f(o)
        return Lazy(self.pool, lambda h: a(h) / b(h))

    def __neg__(self):
        a = self.fn
        return Lazy(self.pool, lambda h: -a(h))


class ScalarPool:
    """Device vector collecting the scalar results of one evaluation (log-determinants, quadratic forms, the per-parameter
    contractions of the gradient).  Every reduction writes its result to its slot on the device (``spde_*_dev``); the
    host reads the vector once.  Replaces one device-to-host copy + stream synchronise per scalar."""

    def __init__(self, cap: int = 2048):
        self.buf = torch.zeros(cap, dtype=F64, device=_dev())
        self.cap, self.used, self.host = cap, 0, None

    def take(self, k: int = 1) -> int:
        i = self.used
        self.used += k
        if self.used > self.cap:
            raise _lib.SpdeError("ScalarPool: more than %d scalars in one evaluation" % self.cap)
        return i

    def addr(self, i: int) -> int:
        return self.buf.data_ptr() + 8 * i

    def scalar(self, i: int) -> Lazy:
        return Lazy(self, lambda h: h[i])

    def fetch(self) -> np.ndarray:
        if self.host is None:
            self.host = to_host(self.buf[:max(self.used, 1)])
        return self.host

    # reductions into the pool (same kernels as Engine.dot / wdot / residual_ss / logdet / gemv_t)
    def dot(self, X, Y) -> Lazy:
        i = self.take()
        check(lib.spde_dot_dev(ptr(X), ptr(Y), X.numel(), self.addr(i), _stream()))
        return self.scalar(i)

    def wdot(self, X, Y, w) -> Lazy:
        i = self.take()
        check(lib.spde_wdot_dev(ptr(X), ptr(Y), ptr(w), X.shape[0], X.shape[1], self.addr(i), _stream()))
        return self.scalar(i)

    def residual_ss(self, data, mu, obs) -> Lazy:
        i = self.take()
        check(lib.spde_residual_ss_dev(ptr(data), ptr(mu), ptr(obs), data.shape[0], data.shape[1], self.addr(i), _stream()))
        return self.scalar(i)

    def logdet(self, engine, which: int) -> Lazy:
        i = self.take()
        check(lib.spde_logdet_dev(engine.plan.h, which, self.addr(i), _stream()))
        return self.scalar(i)

    def gemv_t(self, B, u) -> list:
        """B^T u (B row-major rows x cols) as a list of cols pool scalars."""
        cols = B.shape[1]
        i = self.take(cols)
        check(lib.spde_gemv_t(ptr(B), ptr(u), B.shape[0], cols, self.addr(i), _stream()))
        return [self.scalar(i + c) for c in range(cols)]


class Factor:
    """Drop-in for the ``sksparse.cholmod.Factor`` methods the reference uses
    (``advection_diffusion2D.py:194-202``, ``model.py:80,126``): ``logdet, solve_A, solve_Lt,
    solve_L, apply_P, apply_Pt, P``.  NumPy in -> NumPy out, torch CUDA in -> torch CUDA out.
    The factor lives in store ``which`` of the engine's plan and stays valid until that store is
    refactorised."""

    def __init__(self, engine: "Engine", which: int):
        self.engine = engine
        self.which = which
        self.serial = engine.serial[which]

    def _alive(self):
        if self.engine.serial[self.which] != self.serial:
            raise _lib.SpdeError("this factor has been overwritten by a later factorisation")

    def P(self) -> np.ndarray:
        return self.engine.plan.perm.copy()

    def logdet(self) -> float:
        self._alive()
        return self.engine.logdet(self.which)

    def _solve(self, b, mode):
        self._alive()
        is_np = not isinstance(b, torch.Tensor)
        if is_np and hasattr(b, "toarray"):
            b = b.toarray()
        x = to_dev(np.asarray(b, dtype=np.float64) if is_np else b).clone()
        one_d = x.dim() == 1
        x = x.reshape(self.engine.n, -1).contiguous()
        self.engine.solve(self.which, x, mode)
        if one_d:
            x = x.reshape(-1)
        return x.cpu().numpy() if is_np else x

    def solve_A(self, b):
        return self._solve(b, 15)

    __call__ = solve_A

    def solve_Lt(self, b, use_LDLt_decomposition=False):
        return self._solve(b, 2)

    def solve_L(self, b, use_LDLt_decomposition=False):
        return self._solve(b, 1)

    def apply_P(self, x):
        p = self.engine.plan.perm
        return x[torch.as_tensor(p, device=x.device, dtype=torch.long)] if isinstance(x, torch.Tensor) else np.asarray(x)[p]

    def apply_Pt(self, x):
        p = self.engine.plan.perm
        if isinstance(x, torch.Tensor):
            out = torch.empty_like(x)
            out[torch.as_tensor(p, device=x.device, dtype=torch.long)] = x
            return out
        x = np.asarray(x)
        out = np.empty_like(x)
        out[p] = x
        return out


class Engine:
    _cache: dict = {}

    @classmethod
    def get(cls, M: int, N: int, T: int, bc: int, pat: int = 0) -> "Engine":
        key = (M, N, T, bc, pat, torch.cuda.current_device() if torch.cuda.is_available() else -1)
        if key not in cls._cache:
            cls._cache[key] = cls(M, N, T, bc, pat)
        return cls._cache[key]

    def __init__(self, M: int, N: int, T: int, bc: int, pat: int = 0):
        self.M, self.N, self.T, self.bc, self.pat = M, N, T, bc, pat
        self.bc_abi = bc | (pat << 8)          # pattern selector of the C ABI (SPDE_PATTERN_KRON)
        self.Ns = M * N
        self.n = self.Ns * T
        self.nslots = 25 if T == 1 else (75 if pat == 1 else 43)
        self._plan = None
        self._pattern = None
        self.serial = [0, 0]

    @property
    def plan(self) -> _lib.PlanHandle:
        if self._plan is None:
            self._plan = _lib.PlanHandle(self.M, self.N, self.T, self.bc, self.pat)
        return self._plan

    @property
    def pattern(self) -> Pattern:
        if self._pattern is None:
            self._pattern = Pattern(self.M, self.N, self.T, self.bc, self.pat)
        return self._pattern

    # ------------------------------------------------------------------ assembly (K2, K3)
    def ah_stencil(self, hx, hy, H: torch.Tensor, face: bool) -> torch.Tensor:
        out = torch.empty(9 * self.Ns, dtype=F64, device=_dev())
        check(lib.spde_ah_stencil(self.M, self.N, self.bc, hx, hy, ptr(H), int(face), ptr(out), _stream()))
        return out

    def aw_stencil(self, hx, hy, G: torch.Tensor, dG, face: bool, diff: int, nan_to_zero: bool) -> torch.Tensor:
        out = torch.empty(9 * self.Ns, dtype=F64, device=_dev())
        check(lib.spde_aw_stencil(self.M, self.N, self.bc, hx, hy, ptr(G), ptr(dG), int(face), diff, int(nan_to_zero),
                                  ptr(out), _stream()))
        return out

    def combine_A(self, flavour: int, V, dt, kappa: torch.Tensor, ah, aw) -> torch.Tensor:
        out = torch.empty(9 * self.Ns, dtype=F64, device=_dev())
        kvar = int(kappa is not None and kappa.numel() > 1)
        check(lib.spde_combine_A(self.Ns, flavour, V, dt, ptr(kappa), kvar, ptr(ah), ptr(aw), ptr(out), _stream()))
        return out

    def atda(self, A9: torch.Tensor, kappa: torch.Tensor, V, mode: int) -> torch.Tensor:
        out = torch.empty(25 * self.Ns, dtype=F64, device=_dev())
        check(lib.spde_atda(self.M, self.N, self.bc, ptr(A9), ptr(kappa), int(kappa.numel() > 1), V, mode, ptr(out), _stream()))
        return out

    def fill_spacetime(self, AtDA, A9, kappa, V, Q0_25, sigma, dt, divide: bool) -> torch.Tensor:
        out = torch.empty(43 * self.n, dtype=F64, device=_dev())
        check(lib.spde_fill_spacetime(self.M, self.N, self.T, self.bc, ptr(AtDA), ptr(A9), ptr(kappa), int(kappa.numel() > 1),
                                      V, ptr(Q0_25), sigma, dt, int(divide), ptr(out), _stream()))
        return out

    def fill_kron(self, Qs25: torch.Tensor, d0: float, d1: float, e: float) -> torch.Tensor:
        """Q = Qt (x) Qs in the 75-slot layout; Qt = tridiag(diagonal (d0, d1, ..., d1, d0), off-diagonal e)."""
        out = torch.empty(75 * self.n, dtype=F64, device=_dev())
        check(lib.spde_fill_kron(self.M, self.N, self.T, self.bc, ptr(Qs25), float(d0), float(d1), float(e), ptr(out), _stream()))
        return out

    def kron_reduce(self, W75: torch.Tensor, d0: float, d1: float, e: float) -> torch.Tensor:
        """Adjoint of :meth:`fill_kron` with respect to Qs (weights on the 25-slot spatial pattern)."""
        out = torch.empty(25 * self.Ns, dtype=F64, device=_dev())
        check(lib.spde_kron_reduce(self.M, self.N, self.T, self.bc, ptr(W75), float(d0), float(d1), float(e), ptr(out), _stream()))
        return out

    # ------------------------------------------------------------------ factor / solve / selinv (K4-K7, K10)
    def factorize(self, which: int, Q: torch.Tensor, cnt=None, tau: float = 0.0) -> Factor:
        self.serial[which] += 1
        check(lib.spde_factorize(self.plan.h, which, ptr(Q), ptr(cnt), float(tau), _stream()))
        return Factor(self, which)

    def factorize_async(self, which: int, Q: torch.Tensor, cnt=None, tau: float = 0.0) -> None:
        """Enqueue a factorisation without waiting; pair with :meth:`factor_wait` (lets Q and Q_c overlap)."""
        self.serial[which] += 1
        check(lib.spde_factorize_async(self.plan.h, which, ptr(Q), ptr(cnt), float(tau), _stream()))

    def factor_wait(self, which: int) -> Factor:
        COUNTERS["d2h"] += 4          # the positive-definiteness status word
        check(lib.spde_factor_wait(self.plan.h, which, _stream()))
        return Factor(self, which)

    def selinv_pair(self):
        """Takahashi selected inverse of both stores, the two schedules overlapping on their own streams."""
        check(lib.spde_selinv_start(self.plan.h, 0, _stream()))
        check(lib.spde_selinv_start(self.plan.h, 1, _stream()))
        Z = torch.empty(self.nslots * self.n, dtype=F64, device=_dev())
        Zc = torch.empty(self.nslots * self.n, dtype=F64, device=_dev())
        check(lib.spde_selinv_fetch(self.plan.h, 0, ptr(Z), _stream()))
        check(lib.spde_selinv_fetch(self.plan.h, 1, ptr(Zc), _stream()))
        return Z, Zc

    def selinv_start(self, which: int) -> None:
        """Enqueue the Takahashi pass of a store on its private stream; solves with the same store may run beside
        it (both only read the factor).  Pair with :meth:`selinv_fetch`."""
        check(lib.spde_selinv_start(self.plan.h, which, _stream()))

    def selinv_fetch(self, which: int) -> torch.Tensor:
        Z = torch.empty(self.nslots * self.n, dtype=F64, device=_dev())
        check(lib.spde_selinv_fetch(self.plan.h, which, ptr(Z), _stream()))
        return Z

    def logdet(self, which: int) -> float:
        out = ctypes.c_double()
        check(lib.spde_logdet(self.plan.h, which, ctypes.byref(out), _stream()))
        COUNTERS['d2h'] += 8
        return out.value

    def solve(self, which: int, X: torch.Tensor, mode: int = 15) -> torch.Tensor:
        assert X.is_cuda and X.dtype == F64 and X.is_contiguous() and X.shape[0] == self.n
        check(lib.spde_solve(self.plan.h, which, mode, ptr(X), X.shape[1] if X.dim() > 1 else 1, _stream()))
        return X

    def selinv(self, which: int) -> torch.Tensor:
        Z = torch.empty(self.nslots * self.n, dtype=F64, device=_dev())
        check(lib.spde_selinv(self.plan.h, which, ptr(Z), _stream()))
        return Z

    # ------------------------------------------------------------------ streamed evaluation (csrc/ooc.cu)
    streamed = None      # None: automatic (in-core stores would not fit the device); True / False: forced

    def incore_bytes(self) -> int:
        """Device bytes of one factor store with its update-matrix arenas and inverse fronts (the in-core path)."""
        st = self.plan.stats()
        return st["factor_bytes"] + st["arena_bytes"] + st["zarena_bytes"] + st["dinv_bytes"] + st["ybuf_bytes"]

    def use_streamed(self, stores: int = 1) -> bool:
        """``stores``: factor stores the caller needs resident at once (2 for the Hutchinson mode, which keeps the 3-D
        factors of Q and of Q + tau S^T S; the update-matrix arenas and inverse fronts exist per store as well)."""
        env = os.environ.get("SPDE_STREAMED")
        if env is not None:
            return env not in ("0", "")
        if self.streamed is not None:
            return bool(self.streamed)
        total = torch.cuda.get_device_properties(torch.cuda.current_device()).total_memory
        return stores * self.incore_bytes() + 3 * 8 * self.nslots * self.n > 0.9 * total

    def ooc(self, backward: bool = True) -> _lib.OocHandle:
        """The streamed evaluator of this mesh (one at a time: its pool is most of the device).  The threshold that
        separates front-by-front supernodes (panels parked in pinned host memory) from recomputed subtrees is
        ``SPDE_OOC_TOP_BYTES`` or, by default, the smallest one whose pool and host pool fit this box."""
        cur = getattr(self, "_ooc", None)
        if cur is not None and cur.backward == backward:
            return cur
        self._ooc = None
        del cur
        torch.cuda.empty_cache()
        env = os.environ.get("SPDE_OOC_TOP_BYTES")
        if env is not None:
            thr = int(float(env))
        else:
            import psutil
            total = torch.cuda.get_device_properties(torch.cuda.current_device()).total_memory
            dev_budget = total - 4 * 8 * self.nslots * self.n - (6 << 30)
            host_budget = 0.6 * psutil.virtual_memory().total
            thr = None
            for c in [2.0 ** e * 1e9 for e in range(0, 8)]:      # >= 1 GB: smaller front-by-front segments do not fill the GPU
                st = _lib.OocHandle(self.plan, int(c), backward, False).stats()
                if st["pool_bytes"] <= dev_budget and st["host_bytes"] <= host_budget:
                    thr = int(c)
                    break
            if thr is None:
                raise MemoryError("streamed evaluation of the %dx%dx%d mesh fits neither this device nor this host"
                                  % (self.M, self.N, self.T))
        self._ooc = _lib.OocHandle(self.plan, thr, backward, True)
        return self._ooc

    def streamed_eval(self, Q: torch.Tensor, cnt=None, tau: float = 0.0, X=None, mode: int = 15, selinv: bool = True):
        """One depth-first pass: returns (logdet(Q + tau diag(cnt)), X solved in place or None, Z on the pattern
        of Q or None).  A backward pass (back substitution, selected inverse) runs when asked for."""
        k = 0 if X is None else (X.shape[1] if X.dim() > 1 else 1)
        backward = selinv or (k > 0 and bool(mode & 2))
        o = self.ooc(backward)
        Z = torch.empty(self.nslots * self.n, dtype=F64, device=_dev()) if selinv else None
        if X is not None:
            assert X.is_cuda and X.dtype == F64 and X.is_contiguous() and X.shape[0] == self.n
        ld = o.run(ptr(Q), ptr(cnt), tau, ptr(X), k, mode, ptr(Z), _stream())
        COUNTERS["d2h"] += 8
        return ld, X, Z

    # ------------------------------------------------------------------ reductions (K8, K9, K11)
    def q_apply(self, Q: torch.Tensor, X: torch.Tensor) -> torch.Tensor:
        Y = torch.empty_like(X)
        check(lib.spde_q_apply(self.M, self.N, self.T, self.bc_abi, ptr(Q), ptr(X), X.shape[1], ptr(Y), _stream()))
        return Y

    @staticmethod
    def dot(X: torch.Tensor, Y: torch.Tensor) -> float:
        out = ctypes.c_double()
        check(lib.spde_dot(ptr(X), ptr(Y), X.numel(), ctypes.byref(out), _stream()))
        COUNTERS["d2h"] += 8
        return out.value

    @staticmethod
    def wdot(X: torch.Tensor, Y: torch.Tensor, w: torch.Tensor) -> float:
        out = ctypes.c_double()
        check(lib.spde_wdot(ptr(X), ptr(Y), ptr(w), X.shape[0], X.shape[1], ctypes.byref(out), _stream()))
        COUNTERS["d2h"] += 8
        return out.value

    @staticmethod
    def residual_ss(data: torch.Tensor, mu: torch.Tensor, obs: torch.Tensor) -> float:
        out = ctypes.c_double()
        check(lib.spde_residual_ss(ptr(data), ptr(mu), ptr(obs), data.shape[0], data.shape[1], ctypes.byref(out), _stream()))
        COUNTERS["d2h"] += 8
        return out.value

    def scatter_obs(self, data: torch.Tensor, obs: torch.Tensor, tau: float) -> torch.Tensor:
        b = torch.zeros(self.n, data.shape[1], dtype=F64, device=_dev())
        check(lib.spde_scatter_obs(ptr(data), ptr(obs), data.shape[0], data.shape[1], float(tau), ptr(b), _stream()))
        return b

    def add_diag(self, Q: torch.Tensor, cnt: torch.Tensor, tau: float) -> None:
        diag = Q[(self.nslots // 2) * self.n:(self.nslots // 2 + 1) * self.n]
        check(lib.spde_add_diag(ptr(diag), ptr(cnt), float(tau), self.n, _stream()))

    def sddmm(self, X: torch.Tensor, Y: torch.Tensor, alpha: float, W=None) -> torch.Tensor:
        acc = W is not None
        if W is None:
            W = torch.zeros(self.nslots * self.n, dtype=F64, device=_dev())
        check(lib.spde_sddmm(self.M, self.N, self.T, self.bc_abi, ptr(X), ptr(Y), X.shape[1], float(alpha), int(acc), ptr(W), _stream()))
        return W

    def assembly_adjoint(self, W, A9, kappa, V, sigma, dt, timed: bool):
        Ns = self.Ns
        GA = torch.empty(9 * Ns, dtype=F64, device=_dev())
        if timed:
            work = torch.empty(44 * Ns, dtype=F64, device=_dev())
            Gq = torch.empty(Ns, dtype=F64, device=_dev())
            GQ0 = torch.empty(25 * Ns, dtype=F64, device=_dev())
        else:
            work = Gq = GQ0 = None
        check(lib.spde_assembly_adjoint(self.M, self.N, self.T if timed else 1, self.bc, ptr(W), ptr(A9), ptr(kappa),
                                        int(kappa.numel() > 1), V, sigma, dt, int(timed), ptr(work), ptr(GA), ptr(Gq),
                                        ptr(GQ0), _stream()))
        return GA, Gq, GQ0

    def assembly_adjoint_B(self, WB, A9, kappa, V):
        """Adjoint of ``B = A^T (Qs/V^2) A`` alone (``spde_atda`` mode 1): weights on the Q25 pattern of B ->
        d sum(WB .* B) / d A9 and / d Qs.  Used for the time-collapsed prior log-determinant."""
        Ns = self.Ns
        GA = torch.empty(9 * Ns, dtype=F64, device=_dev())
        Gq = torch.empty(Ns, dtype=F64, device=_dev())
        work = torch.empty(19 * Ns, dtype=F64, device=_dev())
        check(lib.spde_assembly_adjoint(self.M, self.N, 1, self.bc, ptr(WB), ptr(A9), ptr(kappa), int(kappa.numel() > 1),
                                        V, 1.0, 1.0, 2, ptr(work), ptr(GA), ptr(Gq), None, _stream()))
        return GA, Gq

    def stencil_adjoint(self, hx, hy, GA: torch.Tensor, want_H: bool, G=None):
        GH = torch.empty(8 * self.Ns, dtype=F64, device=_dev()) if want_H else None
        GdG = torch.empty(self.Ns, 4, dtype=F64, device=_dev()) if G is not None else None
        check(lib.spde_stencil_adjoint(self.M, self.N, self.bc, hx, hy, ptr(GA), ptr(GH), ptr(G), ptr(GdG), _stream()))
        return GH, GdG

    @staticmethod
    def gemv_t(B: torch.Tensor, u: torch.Tensor) -> torch.Tensor:
        out = torch.empty(B.shape[1], dtype=F64, device=_dev())
        check(lib.spde_gemv_t(ptr(B), ptr(u), B.shape[0], B.shape[1], ptr(out), _stream()))
        return out

    # ------------------------------------------------------------------ export
    def to_scipy(self, Q: torch.Tensor):
        """slot layout -> canonical SciPy CSC (what the reference exposes as ``mod.Q``); exact zeros
        are dropped, as SciPy's SpGEMM / sparse add do in the reference."""
        m = self.pattern.to_csc(to_host(Q))
        m.eliminate_zeros()
        return m
