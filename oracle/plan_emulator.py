"""NumPy interpreter of the device schedules.  TEST INFRASTRUCTURE ONLY.

Executes, launch by launch and tile by tile, the task lists that ``libspde_b200.so`` builds on the
host for the GPU (``spde_plan_export``): the supernodal multifrontal factorisation, the forward /
backward triangular solves and the Takahashi selected inverse.  It exists so that the *plan*
(storage layout, relative indices, tile tables, step ordering) can be validated in the CPU-only
container; it is orders of magnitude too slow to be anything else, and nothing under
``spdepy_b200/`` imports it.  Each handler mirrors the semantics of one CUDA kernel in
``spdepy_b200/csrc/{gemm.cuh,plan.cu}``.
"""
from __future__ import annotations

import numpy as np

GF_LOWER, GF_BETA0, GF_NEG, GF_ATOMIC, GF_GATHER_A, GF_SCATTER_C, GF_MIRROR = (1 << 9, 1 << 10, 1 << 11, 1 << 12,
                                                                            1 << 13, 1 << 14, 1 << 15)
LK_GEMM, LK_POTRF, LK_EXTADD, LK_ZERO, LK_GATHER, LK_WTW, LK_EXTRACT, LK_GEMV, LK_SYNC, LK_COPY, LK_BCOPY = range(11)
NB = 64
CFG = {0: (128, 128), 1: (128, 64), 2: (64, 64), 3: (128, 64)}      # 3 = warp-specialised bulk-async kernel

LAUNCH = np.dtype([("kind", "i4"), ("variant", "i4"), ("task0", "i8"), ("ntasks", "i4"), ("tile0", "i8"),
                   ("ntiles", "i4"), ("a0", "i8"), ("a1", "i8"), ("lane", "i4"), ("pad", "i4")], align=True)
GEMM = np.dtype([("a", "i8"), ("b", "i8"), ("c", "i8"), ("c2", "i8"), ("lda", "i4"), ("ldb", "i4"), ("ldc", "i4"),
                 ("M", "i4"), ("N", "i4"), ("K", "i4"), ("flags", "i4"), ("aidx", "i4"), ("cidx", "i4"), ("pad", "i4")],
                align=True)
TILE = np.dtype([("task", "i4"), ("ti", "i4"), ("tj", "i4"), ("pad", "i4")], align=True)
POTRF = np.dtype([("blk", "i8"), ("dinv", "i8"), ("ld", "i4"), ("b", "i4"), ("col0", "i4"), ("pad", "i4")], align=True)
EXT = np.dtype([("src", "i8"), ("lds", "i4"), ("nr", "i4"), ("ppanel", "i8"), ("pld", "i4"), ("pnc", "i4"),
                ("pncp", "i4"), ("pupd", "i8"), ("pldu", "i4"), ("rel", "i4"), ("src_space", "i4"),
                ("dst_space", "i4")], align=True)
GATHER = np.dtype([("dst", "i8"), ("ldd", "i4"), ("ncp", "i4"), ("nr", "i4"), ("src", "i8"), ("lds", "i4"),
                   ("pnc", "i4"), ("pncp", "i4"), ("rel", "i4"), ("src_space", "i4"), ("dst_space", "i4")], align=True)
WTW = np.dtype([("w", "i8"), ("dst", "i8"), ("ldd", "i4"), ("b", "i4"), ("space", "i4"), ("pad", "i4")], align=True)
BCOPY = np.dtype([("dst", "i8"), ("src", "i8"), ("ldd", "i4"), ("lds", "i4"), ("rows", "i4"), ("cols", "i4"),
                  ("dst_space", "i4"), ("src_space", "i4")], align=True)
ZENT = np.dtype([("dst", "i8"), ("dst2", "i8"), ("src", "i8"), ("sn", "i4"), ("pad", "i4")], align=True)


SCAT = np.dtype([("src", "i8"), ("dst", "i8")], align=True)


class Program:
    def __init__(self, plan, prog, k=0, seg=None):
        if seg is None:
            ex = plan.export
        else:       # schedule of one segment of the streamed evaluator (spde_ooc_export)
            def ex(prog, what, dtype, k):
                return plan.export(seg, prog, what, dtype, k)
        self.launches = ex(prog, 0, LAUNCH, k)
        self.gemm = ex(prog, 1, GEMM, k)
        self.tiles = ex(prog, 2, TILE, k)
        self.potrf = ex(prog, 3, POTRF, k)
        self.ext = ex(prog, 4, EXT, k)
        self.gather = ex(prog, 5, GATHER, k)
        self.wtw = ex(prog, 6, WTW, k)
        try:
            self.bcopy = ex(prog, 8, BCOPY, k)
        except Exception:       # (the streamed evaluator's export has no block-copy table)
            self.bcopy = np.zeros(0, BCOPY)


class Emulator:
    def __init__(self, plan):
        self.plan = plan
        self.n = plan.n
        self.nslots = plan.nslots
        sizes = plan.export(4, 0, "i8")
        (self.l_size, self.dinv_size, a0, a1, z0, z1, self.ybuf_size, self.rel_base) = [int(v) for v in sizes[:8]]
        self.solve_outer = bool(sizes[8]) if len(sizes) > 8 else False      # solves ping-pong between X and X2 (space 1)
        self.qdest = plan.export(4, 1, "i8")
        self.cand = plan.export(4, 2, "i4")
        self.diagpos = plan.export(4, 3, "i8")
        self.idx = plan.export(4, 4, "i4")
        self.perm = plan.perm.astype(np.int64)
        self.sp = [None] * 8
        self.sp[0] = np.zeros(max(self.l_size, 2))
        self.sp[1] = np.zeros(max(a0, 1))
        self.sp[2] = np.zeros(max(a1, 1))
        self.sp[3] = np.zeros(max(self.dinv_size, 2))
        self.sp[5] = np.zeros(max(self.ybuf_size, 2))      # scratch of the outer-block products (factor and Takahashi)
        self.zsizes = (z0, z1)
        self.status = 0

    # ---- kernels -----------------------------------------------------------------------------
    def _view(self, space, off, ld, rows, cols):
        """column-major (rows x cols) view with leading dimension ld"""
        buf = self.sp[space]
        return np.lib.stride_tricks.as_strided(buf[off:], shape=(rows, cols), strides=(8, 8 * ld), writeable=True)

    def _gemm_launch(self, P, L, gemv=False):
        if gemv:    # k_gemv_grouped: all M rows, 256 / 64 columns per tile, K split in chunks of 1024 (tile.ti)
            ak, bk = 0, int(L["variant"])
            BM, BN = 1 << 30, (64 if bk else 256)
        else:
            cfg, ak, bk = L["variant"] // 4, (L["variant"] >> 1) & 1, L["variant"] & 1
            BM, BN = CFG[cfg]
        tasks = P.gemm[L["task0"]:L["task0"] + L["ntasks"]]
        tiles = P.tiles[L["tile0"]:L["tile0"] + L["ntiles"]]
        # the CUDA kernel reads its operand tiles completely before writing C; emulate per tile
        for tr in tiles:
            t = tasks[tr["task"]]
            f = int(t["flags"])
            M, N, K = int(t["M"]), int(t["N"]), int(t["K"])
            if gemv:
                assert M <= 4
                i0, j0 = 0, int(tr["tj"]) * BN
                kc = 1024 if bk else 256
                ka, kb = int(tr["ti"]) * kc, min(K, int(tr["ti"]) * kc + kc)
                assert not (f & GF_BETA0) or K <= kc
            else:
                i0, j0 = int(tr["ti"]) * BM, int(tr["tj"]) * BN
                ka, kb = 0, K
            i1, j1 = min(i0 + BM, M), min(j0 + BN, N)
            sa, sb, sc = f & 7, (f >> 3) & 7, (f >> 6) & 7
            if ak:
                A = self._view(sa, int(t["a"]), int(t["lda"]), K, M).T[i0:i1]
            elif f & GF_GATHER_A:
                cols = self.idx[int(t["aidx"]):int(t["aidx"]) + K].astype(np.int64)
                full = self.sp[sa]
                A = np.stack([full[int(t["a"]) + c * int(t["lda"]) + i0: int(t["a"]) + c * int(t["lda"]) + i1] for c in cols], axis=1)
            else:
                A = self._view(sa, int(t["a"]), int(t["lda"]), M, K)[i0:i1]
            if bk:
                B = self._view(sb, int(t["b"]), int(t["ldb"]), K, N)[:, j0:j1]
            else:
                B = self._view(sb, int(t["b"]), int(t["ldb"]), N, K)[j0:j1].T
            prod = np.array(A)[:, ka:kb] @ np.array(B)[ka:kb]
            if f & GF_NEG:
                prod = -prod
            ii, jj = np.meshgrid(np.arange(i0, i1), np.arange(j0, j1), indexing="ij")
            keep = (ii >= jj) if (f & GF_LOWER) else np.ones_like(ii, dtype=bool)
            ldc = int(t["ldc"])
            if f & GF_SCATTER_C:
                cmap = self.idx[int(t["cidx"]):int(t["cidx"]) + N].astype(np.int64)
                cols = cmap[jj]
            else:
                cols = jj
            lin = int(t["c"]) + ii + cols * ldc
            buf = self.sp[sc]
            if f & GF_BETA0:
                buf[lin[keep]] = prod[keep]
            else:
                np.add.at(buf, lin[keep], prod[keep])
            if f & GF_MIRROR:
                lin2 = int(t["c2"]) + jj + ii * ldc
                if f & GF_BETA0:
                    buf[lin2[keep]] = prod[keep]
                else:
                    np.add.at(buf, lin2[keep], prod[keep])

    def _potrf(self, P, L):
        for t in P.potrf[L["task0"]:L["task0"] + L["ntasks"]]:
            b, ld = int(t["b"]), int(t["ld"])
            blk = self._view(0, int(t["blk"]), ld, b, b)
            A = np.tril(np.array(blk))
            A = A + np.tril(A, -1).T
            try:
                Lc = np.linalg.cholesky(A)
            except np.linalg.LinAlgError:
                self.status = int(t["col0"]) + 1
                Lc = np.full((b, b), np.nan)
            blk[:, :] = Lc
            W = np.zeros((NB, NB))
            if self.status == 0:
                W[:b, :b] = np.linalg.solve(Lc, np.eye(b))
                W[:b, :b] = np.tril(W[:b, :b])
            self.sp[3][int(t["dinv"]):int(t["dinv"]) + NB * NB] = W.flatten(order="F")

    def _extadd(self, P, L):
        tasks = P.ext[L["task0"]:L["task0"] + L["ntasks"]]
        tiles = P.tiles[L["tile0"]:L["tile0"] + L["ntiles"]]
        for tr in tiles:
            t = tasks[tr["task"]]
            nr = int(t["nr"])
            rel = self.idx[int(t["rel"]):int(t["rel"]) + nr].astype(np.int64)
            src = self._view(int(t["src_space"]), int(t["src"]), int(t["lds"]), nr, nr)
            i = np.arange(int(tr["ti"]) * 32, min(int(tr["ti"]) * 32 + 32, nr))
            j = np.arange(int(tr["tj"]) * 32, min(int(tr["tj"]) * 32 + 32, nr))
            ii, jj = np.meshgrid(i, j, indexing="ij")
            keep = jj <= ii
            ii, jj = ii[keep], jj[keep]
            v = src[ii, jj]
            ri, rj = rel[ii], rel[jj]
            pnc, pncp = int(t["pnc"]), int(t["pncp"])
            inpanel = rj < pnc
            row = np.where(ri < pnc, ri, pncp + (ri - pnc))
            np.add.at(self.sp[0], int(t["ppanel"]) + row[inpanel] + rj[inpanel] * int(t["pld"]), v[inpanel])
            o = ~inpanel
            np.add.at(self.sp[int(t["dst_space"])], int(t["pupd"]) + (ri[o] - pnc) + (rj[o] - pnc) * int(t["pldu"]), v[o])

    def _gather(self, P, L):
        tasks = P.gather[L["task0"]:L["task0"] + L["ntasks"]]
        tiles = P.tiles[L["tile0"]:L["tile0"] + L["ntiles"]]
        for tr in tiles:
            t = tasks[tr["task"]]
            nr = int(t["nr"])
            rel = self.idx[int(t["rel"]):int(t["rel"]) + nr].astype(np.int64)
            pnc, pncp = int(t["pnc"]), int(t["pncp"])
            rel = np.where(rel < pnc, rel, pncp + (rel - pnc))
            i = np.arange(int(tr["ti"]) * 32, min(int(tr["ti"]) * 32 + 32, nr))
            j = np.arange(int(tr["tj"]) * 32, min(int(tr["tj"]) * 32 + 32, nr))
            ii, jj = np.meshgrid(i, j, indexing="ij")
            src = self.sp[int(t["src_space"])]
            dst = self.sp[int(t["dst_space"])]
            ncp, ldd = int(t["ncp"]), int(t["ldd"])
            vals = src[int(t["src"]) + rel[ii] + rel[jj] * int(t["lds"])]
            dst[int(t["dst"]) + (ncp + ii) + (ncp + jj) * ldd] = vals
            if int(tr["ti"]) != int(tr["tj"]):      # k_selinv_gather reads the tiles on and below the diagonal and mirrors them
                dst[int(t["dst"]) + (ncp + jj) + (ncp + ii) * ldd] = vals

    def _wtw(self, P, L):
        for t in P.wtw[L["task0"]:L["task0"] + L["ntasks"]]:
            b = int(t["b"])
            if int(t["pad"]) == 1:      # copy mode: the 64x64 inverses of an outer block onto the block diagonal of Wf
                for j in range((b + NB - 1) // NB):
                    bj = min(NB, b - j * NB)
                    Wj = self.sp[3][int(t["w"]) + j * NB * NB:int(t["w"]) + (j + 1) * NB * NB].reshape(NB, NB, order="F")[:bj, :bj]
                    self._view(int(t["space"]), int(t["dst"]) + j * NB * (int(t["ldd"]) + 1), int(t["ldd"]), bj, bj)[:, :] = Wj
                continue
            W = self.sp[3][int(t["w"]):int(t["w"]) + NB * NB].reshape(NB, NB, order="F")[:b, :b]
            self._view(int(t["space"]), int(t["dst"]), int(t["ldd"]), b, b)[:, :] = W.T @ W

    def run(self, P, Zq=None, zent=None):
        for L in P.launches:
            kind = int(L["kind"])
            if kind == LK_GEMM:
                self._gemm_launch(P, L)
            elif kind == LK_GEMV:
                self._gemm_launch(P, L, gemv=True)
            elif kind == LK_POTRF:
                self._potrf(P, L)
            elif kind == LK_EXTADD:
                self._extadd(P, L)
            elif kind == LK_ZERO:
                self.sp[int(L["variant"])][int(L["a0"]):int(L["a1"])] = 0.0
            elif kind == LK_GATHER:
                self._gather(P, L)
            elif kind == LK_WTW:
                self._wtw(P, L)
            elif kind == LK_BCOPY:
                for t in P.bcopy[L["task0"]:L["task0"] + L["ntasks"]]:
                    src = np.array(self._view(int(t["src_space"]), int(t["src"]), int(t["lds"]), int(t["rows"]), int(t["cols"])))
                    self._view(int(t["dst_space"]), int(t["dst"]), int(t["ldd"]), int(t["rows"]), int(t["cols"]))[:, :] = src
            elif kind == LK_SYNC:
                continue        # lane ordering: the launch list is a valid serial order
            elif kind == LK_COPY:
                self._copy(L)
            elif kind == LK_EXTRACT:
                e = zent[int(L["a0"]):int(L["a1"])]
                v = self.sp[int(L["variant"])][e["src"]]
                Zq[e["dst"]] = v
                m = e["dst2"] >= 0
                Zq[e["dst2"][m]] = v[m]

    def _copy(self, L):
        raise NotImplementedError("LK_COPY only occurs in the schedules of the streamed evaluator")

    # ---- entry points mirroring the C ABI -------------------------------------------------------
    def factorize(self, Qslots, cnt=None, tau=0.0):
        """``spde_factorize``: Qslots is the flat slot-major array (nslots*n)."""
        self.sp[0][:] = 0.0
        n = self.n
        for ci, slot in enumerate(self.cand):
            d = self.qdest[ci * n:(ci + 1) * n]
            m = d >= 0
            v = Qslots[slot * n:(slot + 1) * n].copy()
            if cnt is not None and slot == self.nslots // 2:
                v = v + cnt * tau
            self.sp[0][d[m]] = v[m]
        self.status = 0
        self.run(Program(self.plan, 0))
        return self.status

    def logdet(self):
        return 2.0 * np.log(self.sp[0][self.diagpos]).sum()

    def solve(self, X, mode=15):
        """``spde_solve`` on a row-major (n,k) array (returns a new array)."""
        X = np.asarray(X, dtype=np.float64)
        X = X.reshape(self.n, -1)
        k = X.shape[1]
        kp = k + (k & 1)
        Xp = np.zeros((self.n, kp))
        Xp[:, :k] = X[self.perm] if mode & 4 else X
        arena0 = self.sp[1]
        if self.solve_outer:
            # spde_solve: space 1 is the second right-hand-side buffer; forward reads X and leaves y in X2, backward reads
            # y from X2 and leaves x in X (stale NaNs where the schedule must not read)
            self.sp[4] = np.full(self.n * kp, np.nan)
            self.sp[1] = np.full(self.n * kp, np.nan)
            self.sp[4 if mode & 1 else 1] = Xp.reshape(-1).copy()
        else:
            self.sp[4] = Xp.reshape(-1).copy()
        if mode & 1:
            self.run(Program(self.plan, 1, k))
        if mode & 2:
            self.run(Program(self.plan, 2, k))
        res = self.sp[1 if (self.solve_outer and not mode & 2) else 4]
        self.sp[1] = arena0
        Xp = res.reshape(self.n, kp)[:, :k]
        out = np.empty_like(Xp)
        if mode & 8:
            out[self.perm] = Xp
        else:
            out[:] = Xp
        return out

    def selinv(self):
        """``spde_selinv``: returns Z on the pattern of Q, flat slot-major (nslots*n)."""
        self.sp[6] = np.zeros(max(self.zsizes[0], 1))
        self.sp[7] = np.zeros(max(self.zsizes[1], 1))
        zent = self.plan.export(4, 5, ZENT)
        Zq = np.zeros(self.nslots * self.n)
        self.run(Program(self.plan, 3), Zq=Zq, zent=zent)
        return Zq


class OocEmulator(Emulator):
    """Interpreter of the streamed evaluator (``csrc/ooc.cu``, ``spde_ooc_run``): one pool array aliased by all
    operand spaces, segments run depth-first in the exported order, panels of top segments parked in a host array
    between the passes, bottom segments factorised again in the backward pass."""

    def __init__(self, plan, ooc):
        self.plan, self.ooc = plan, ooc
        self.n, self.nslots = plan.n, plan.nslots
        self.perm = plan.perm.astype(np.int64)
        tab = ooc.export(-1, 0, 0, "i8").reshape(-1, 16)
        names = ("top", "keep", "root", "parent", "off_L", "l_size", "off_dinv", "dinv_size", "upd", "u_size", "stack_U",
                 "stack_Z", "scat0", "scat1", "col0", "col1")
        self.segs = [dict(zip(names, [int(v) for v in row])) for row in tab]
        self.order = ooc.export(-1, 0, 1, "i4")
        self.scat = ooc.export(-1, 0, 2, SCAT)
        self.diagpos = ooc.export(-1, 0, 3, "i8")
        self.idx = ooc.export(-1, 0, 4, "i4")
        self.zent = ooc.export(-1, 0, 5, ZENT)
        self.zent0 = ooc.export(-1, 0, 6, "i8")
        self.pool_size = ooc.info(2) // 8
        self.pool = np.full(self.pool_size, np.nan)      # stale memory must never be read as zeros
        self.host = {}
        self.hostbuf = None      # pinned host pool (slices of the overlapped top segments)
        self.chunks = None       # slice table of the segment in flight
        self.sp = [self.pool] * 8
        self.status = 0

    def _copy(self, L):
        """LK_COPY: variant 0 parks pool[a0:a0+a1] at host offset task0 (only when a backward pass follows);
        variant 1 is the point where slice a0 of the panel (and every later slice) must have come back -- the
        interpreter fetches them exactly there, so a step that reads a slice before its wait record sees the NaN poison."""
        if int(L["variant"]) == 0:
            if self.hostbuf is not None:
                a0, a1, h = int(L["a0"]), int(L["a1"]), int(L["task0"])
                self.hostbuf[h:h + a1] = self.pool[a0:a0 + a1]
        else:
            # (the slices come back last slice first on ONE stream: when slice a0 has arrived, so have all later ones)
            for c in range(int(L["a0"]), len(self.chunks)):
                off, ln, h = (int(v) for v in self.chunks[c])
                self.pool[off:off + ln] = self.hostbuf[h:h + ln]

    def _scatter_factor(self, s, g, Qslots, cnt, tau):
        self.pool[g["off_L"]:g["off_L"] + g["l_size"]] = 0.0
        e = self.scat[g["scat0"]:g["scat1"]]
        v = Qslots[e["src"]].copy()
        if cnt is not None:
            slot, node = e["src"] // self.n, e["src"] % self.n
            dm = slot == self.nslots // 2
            v[dm] += cnt[node[dm]] * tau
        self.pool[e["dst"]] = v
        self.run(Program(self.ooc, 0, seg=s))

    def evaluate(self, Qslots, cnt=None, tau=0.0, X=None, mode=15, selinv=True):
        """``spde_ooc_run``: returns (logdet, X solved, Zq)."""
        k = 0
        if X is not None:
            X = np.asarray(X, dtype=np.float64).reshape(self.n, -1)
            k = X.shape[1]
            kp = k + (k & 1)
            Xp = np.zeros((self.n, kp))
            Xp[:, :k] = X[self.perm] if mode & 4 else X
            self.sp = list(self.sp)
            self.sp[4] = Xp.reshape(-1).copy()
        fs, bs = k > 0 and bool(mode & 1), k > 0 and bool(mode & 2)
        backward = bs or selinv
        ld = {}
        end = self.pool_size
        self.hostbuf = np.full(max(self.ooc.info(5) // 8, 1), np.nan) if backward else None
        slices = {s: self.ooc.export(s, 0, 7, "i8").reshape(-1, 3) for s in range(len(self.segs))}
        for s in self.order:
            g = self.segs[s]
            self._scatter_factor(s, g, Qslots, cnt, tau)
            ld[s] = 2.0 * np.log(self.pool[self.diagpos[g["col0"]:g["col1"]]]).sum()
            if fs:
                self.run(Program(self.ooc, 1, k, seg=s))
            if backward and g["top"] and not g["keep"] and not len(slices[s]):
                self.host[s] = self.pool[g["off_dinv"]:end].copy()      # (overlapped segments parked their slices already)
            if g["u_size"]:
                self.pool[g["stack_U"]:g["stack_U"] + g["u_size"]] = self.pool[g["upd"]:g["upd"] + g["u_size"]].copy()
            if not (backward and g["keep"]):
                # nothing above the stack may be relied upon later (except the root kept between the passes)
                self.pool[g["stack_U"] + g["u_size"]:end] = np.nan
        Zq = np.zeros(self.nslots * self.n) if selinv else None
        if backward:
            for s in self.order[::-1]:
                g = self.segs[s]
                fetch = False
                if not g["keep"]:
                    if g["top"] and len(slices[s]):
                        fetch = True
                        off, ln, h = (int(v) for v in slices[s][-1])          # inverse diagonal blocks first
                        self.pool[off:off + ln] = self.hostbuf[h:h + ln]
                        self.chunks = slices[s][:-1]
                    elif g["top"]:
                        self.pool[g["off_dinv"]:end] = self.host.pop(s)
                    else:
                        self._scatter_factor(s, g, Qslots, cnt, tau)
                if bs and not fetch:
                    self.run(Program(self.ooc, 2, k, seg=s))
                if selinv:
                    z0 = int(self.zent0[s])
                    z1 = int(self.zent0[s + 1]) if s + 1 < len(self.segs) else len(self.zent)
                    self.run(Program(self.ooc, 3, seg=s), Zq=Zq, zent=self.zent[z0:z1])
                if fetch:
                    if not selinv:       # spde_ooc_run waits for the last slice (number 0) before the solve
                        for off, ln, h in self.chunks:
                            self.pool[int(off):int(off + ln)] = self.hostbuf[int(h):int(h + ln)]
                    if bs:
                        self.run(Program(self.ooc, 2, k, seg=s))
                kids = [c for c in range(len(self.segs)) if self.segs[c]["parent"] == s]
                top = max([self.segs[c]["stack_Z"] + self.segs[c]["u_size"] for c in kids], default=max(g["stack_Z"], 0))
                self.pool[top:end] = np.nan
        out = None
        if k:
            Xp = self.sp[4].reshape(self.n, -1)[:, :k]
            out = np.empty_like(Xp)
            if mode & 8:
                out[self.perm] = Xp
            else:
                out[:] = Xp
        self.last_ld = ld
        return sum(ld[s] for s in self.order), out, Zq
