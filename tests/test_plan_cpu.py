"""CPU validation of the host-side symbolic analysis and of the device schedules.

The schedules built by libspde_b200.so for the GPU (factorisation, solves, Takahashi) are
executed by the NumPy interpreter in oracle/plan_emulator.py and compared with dense LAPACK on the
oracle's Q.  No CUDA call is made here."""
import numpy as np
import pytest
from scipy import sparse

import plan_emulator as pe
import spde_oracle as so
from helpers import load_golden, make_oracle
from spdepy_b200 import _lib
from spdepy_b200.pattern import Pattern

CASES = ["wm_iso_bc3", "wm_ani_bc1_ext", "varwm_ani_bc2", "ad_ani_bc3_q0", "ad_ha_bc1_q0", "vavd_ani_bc2"]


def _setup(name):
    d = load_golden(name)
    mod = make_oracle(d)
    mod.setQ(d["par"])
    M, N = mod.grid.shape[0], mod.grid.shape[1]
    T = mod.grid.T if mod.spec.timed else 1
    plan = _lib.PlanHandle(M, N, T, d["bc"])
    pat = Pattern(M, N, T, d["bc"])
    return d, mod, plan, pat


def _dense_sym(Q):
    A = sparse.tril(Q).toarray()
    return A + np.tril(A, -1).T


@pytest.mark.parametrize("name", CASES)
def test_pattern_roundtrip(name):
    d, mod, plan, pat = _setup(name)
    flat = pat.from_sparse(mod.Q)
    back = pat.to_csc(flat)
    assert abs(back - mod.Q).max() == 0.0


@pytest.mark.parametrize("name", CASES)
def test_symbolic_structure(name):
    d, mod, plan, pat = _setup(name)
    perm = plan.perm.astype(np.int64)
    assert sorted(perm.tolist()) == list(range(plan.n))
    first, rowptr, rows, parent = plan.supernodes()
    # boolean symbolic factorisation of the permuted mesh pattern
    S = (pat.to_csc(np.ones(pat.nslots * pat.n)).toarray() != 0)[np.ix_(perm, perm)]
    n = plan.n
    Lb = np.tril(S)
    for j in range(n):
        r = np.nonzero(Lb[j + 1:, j])[0] + j + 1
        if r.size:
            Lb[np.ix_(r, r)] |= True
    Lb = np.tril(Lb)
    cc = Lb.sum(axis=0)
    assert int(cc.sum()) == plan.info(2)
    assert abs(float((cc.astype(float) ** 2).sum()) - plan.info_d(3)) < 0.5
    # every column's structure must be inside its supernode's (relaxed supernodes may add zeros)
    for s in range(first.size - 1):
        below = set(rows[rowptr[s]:rowptr[s + 1]].tolist())
        for j in range(first[s], first[s + 1]):
            r = np.nonzero(Lb[:, j])[0]
            r = r[r >= first[s + 1]]
            assert set(r.tolist()) <= below


@pytest.mark.parametrize("name", CASES)
def test_emulated_factor_solve_selinv(name):
    d, mod, plan, pat = _setup(name)
    n = plan.n
    flat = pat.from_sparse(mod.Q)
    em = pe.Emulator(plan)
    rng = np.random.default_rng(0)
    cnt = np.zeros(n)
    cnt[rng.choice(n, n // 3, replace=False)] = 1.0
    tau = 7.0
    A = _dense_sym(mod.Q) + np.diag(cnt * tau)
    assert em.factorize(flat, cnt, tau) == 0
    sign, ld = np.linalg.slogdet(A)
    assert abs(em.logdet() - ld) < 1e-10 * abs(ld)
    perm = plan.perm.astype(np.int64)
    Lref = np.linalg.cholesky(A[np.ix_(perm, perm)])
    B = rng.normal(size=(n, 3))
    X = em.solve(B, mode=15)
    assert np.abs(X - np.linalg.solve(A, B)).max() < 1e-9 * np.abs(X).max()
    Zs = em.solve(B, mode=10)        # P^T L^-T z   (model.py:80)
    ref = np.empty_like(B)
    ref[perm] = np.linalg.solve(Lref.T, B)
    assert np.abs(Zs - ref).max() < 1e-9 * np.abs(ref).max()
    Y = em.solve(B, mode=5)         # L^-1 P b
    assert np.abs(Y - np.linalg.solve(Lref, B[perm])).max() < 1e-9 * np.abs(Y).max()
    X1 = em.solve(B[:, 0], mode=15)  # odd number of right-hand sides
    assert np.abs(X1[:, 0] - X[:, 0]).max() < 1e-12 * np.abs(X).max()
    # Takahashi selected inverse against the dense inverse on the pattern of Q
    Zq = em.selinv()
    Zd = np.linalg.inv(A)
    full = pat.to_csc(Zq).toarray()
    mask = pat.to_csc(np.ones(pat.nslots * n)).toarray() != 0
    assert np.abs(full[mask] - Zd[mask]).max() < 1e-9 * np.abs(Zd).max()


def test_not_spd_is_reported():
    d, mod, plan, pat = _setup("wm_iso_bc3")
    flat = pat.from_sparse(mod.Q)
    flat[(pat.nslots // 2) * plan.n + 5] = -1.0     # a negative diagonal entry
    em = pe.Emulator(plan)
    assert em.factorize(flat) != 0


@pytest.mark.parametrize("shape", [(24, 22, 9, 3), (40, 37, 1, 1), (21, 20, 6, 2)])
def test_emulated_large_fronts(shape):
    """Synthetic SPD matrix on the mesh pattern, big enough for multi-block supernodes
    (nc > 256 exercises the outer right-looking blocks)."""
    M, N, T, bc = shape
    plan = _lib.PlanHandle(M, N, T, bc)
    assert plan.info(13) > (256 if T > 1 else 64)
    pat = Pattern(M, N, T, bc)
    n = plan.n
    rng = np.random.default_rng(1)
    W = pat.to_csc(rng.normal(size=pat.nslots * n))
    A = (W + W.T) * 0.5
    A = A + sparse.diags(np.abs(A).sum(axis=1).A1 + 1.0)
    A = sparse.csc_matrix(A)
    flat = pat.from_sparse(A)
    em = pe.Emulator(plan)
    assert em.factorize(flat) == 0
    Ad = A.toarray()
    sign, ld = np.linalg.slogdet(Ad)
    assert abs(em.logdet() - ld) < 1e-11 * abs(ld)
    B = rng.normal(size=(n, 2))
    X = em.solve(B, mode=15)
    assert np.abs(Ad @ X - B).max() < 1e-10 * np.abs(B).max()
    Zq = em.selinv()
    Zd = np.linalg.inv(Ad)
    full = pat.to_csc(Zq).toarray()
    mask = pat.to_csc(np.ones(pat.nslots * n)).toarray() != 0
    assert np.abs(full[mask] - Zd[mask]).max() < 1e-10 * np.abs(Zd).max()


@pytest.mark.parametrize("blocks,kchunk2,diag,solve_outer", [("2", "0", "1", "1"), ("1", "2048", "1", "0"), ("2", "0", "0", "1"),
                                                             ("3", "96", "0", "1"), ("3", "96", "0", "0")])
def test_emulated_selinv_outer_blocks(monkeypatch, blocks, kchunk2, diag, solve_outer):
    """Two-level Takahashi recursion with several outer blocks per front (normally 512 columns wide): outer-block
    inverses by recursive doubling (ragged last block, non-power-of-two block counts), unsplit and K-chunked products;
    and the same mesh through the 64-column recursion (SPDE_SELINV_OUTER=0) gives the same selected inverse.  diag = 1: the
    factorisation itself works on outer blocks (diagonal-first panels, one out-of-place TRSM with the outer-block inverse
    per 512 columns, block copies) and leaves the inverses for the recursion; diag = 0: blocked left-looking panels, the
    recursion builds the inverses (what the streamed schedules do) -- unless solve_outer = 1 (the default): then the factor
    schedule ends with an epilogue that builds the inverses of the whole tree, and the triangular solves work on outer
    blocks with two right-hand-side buffers."""
    monkeypatch.setenv("SPDE_FACTOR_DIAG", diag)
    monkeypatch.setenv("SPDE_SOLVE_OUTER_BLOCKS", solve_outer)
    monkeypatch.setenv("SPDE_SELINV_OUTER_BLOCKS", blocks)
    monkeypatch.setenv("SPDE_SELINV_KCHUNK2", kchunk2)
    M, N, T, bc = 24, 22, 9, 3
    plan = _lib.PlanHandle(M, N, T, bc)
    assert plan.info(13) > 64 * 2 * int(blocks)          # the widest front spans more than two outer blocks
    pat = Pattern(M, N, T, bc)
    n = plan.n
    rng = np.random.default_rng(3)
    W = pat.to_csc(rng.normal(size=pat.nslots * n))
    A = (W + W.T) * 0.5
    A = sparse.csc_matrix(A + sparse.diags(np.abs(A).sum(axis=1).A1 + 1.0))
    em = pe.Emulator(plan)
    assert em.factorize(pat.from_sparse(A)) == 0
    assert em.solve_outer == (solve_outer == "1")
    in_factor = diag == "1" or solve_outer == "1"
    prog = pe.Program(plan, 0 if in_factor else 3)
    assert (prog.wtw["pad"] == 1).any()                    # copy-mode tasks: the outer-block inverses are built here
    assert (len(pe.Program(plan, 0).bcopy) > 0) == (diag == "1")
    Ad = A.toarray()
    sign, ld0 = np.linalg.slogdet(Ad)
    assert abs(em.logdet() - ld0) < 1e-11 * abs(ld0)
    Bm = rng.normal(size=(n, 2))
    assert np.abs(Ad @ em.solve(Bm, mode=15) - Bm).max() < 1e-10 * np.abs(Bm).max()
    perm = plan.perm.astype(np.int64)
    Lref = np.linalg.cholesky(Ad[np.ix_(perm, perm)])
    for k in ((1, 5) if kchunk2 == "0" else (2,)):      # forward only / backward only, matrix-vector and tile kernels, odd column counts
        Bk = rng.normal(size=(n, k))
        Y = em.solve(Bk, mode=5)         # L^-1 P b
        assert np.abs(Y - np.linalg.solve(Lref, Bk[perm])).max() < 1e-10 * np.abs(Y).max()
        Zs = em.solve(Bk, mode=10)       # P^T L^-T z
        ref = np.empty_like(Bk)
        ref[perm] = np.linalg.solve(Lref.T, Bk)
        assert np.abs(Zs - ref).max() < 1e-10 * np.abs(ref).max()
    Zq = em.selinv()
    Zd = np.linalg.inv(A.toarray())
    full = pat.to_csc(Zq).toarray()
    mask = pat.to_csc(np.ones(pat.nslots * n)).toarray() != 0
    assert np.abs(full[mask] - Zd[mask]).max() < 1e-10 * np.abs(Zd).max()
    monkeypatch.setenv("SPDE_SELINV_OUTER", "0")
    plan0 = _lib.PlanHandle(M, N, T, bc)
    em0 = pe.Emulator(plan0)
    assert em0.factorize(pat.from_sparse(A)) == 0
    assert not (pe.Program(plan0, 3).wtw["pad"] == 1).any()
    Z0 = em0.selinv()
    assert np.abs(Z0 - Zq).max() < 1e-11 * np.abs(Zd).max()


def test_emulated_tail_split(monkeypatch):
    """Tail split of many-wave launches (the tiles of the partly filled last wave become one-tile, K-chunked, atomically
    accumulating tasks): with a pretended machine of 6 slots it triggers on small meshes, in the factorisation (lower-masked
    updates, tiles on the diagonal are kept whole), the solves and the Takahashi products (mirror writes, zero-destination)."""
    monkeypatch.setenv("SPDE_TAIL_SLOTS", "6")
    M, N, T, bc = 24, 22, 9, 3
    plan = _lib.PlanHandle(M, N, T, bc)
    pat = Pattern(M, N, T, bc)
    n = plan.n
    rng = np.random.default_rng(6)
    W = pat.to_csc(rng.normal(size=pat.nslots * n))
    A = (W + W.T) * 0.5
    A = sparse.csc_matrix(A + sparse.diags(np.abs(A).sum(axis=1).A1 + 1.0))
    em = pe.Emulator(plan)
    assert em.factorize(pat.from_sparse(A)) == 0
    for prog in (0, 3):
        P = pe.Program(plan, prog)
        one_tile = sum(1 for L in P.launches if L["kind"] == pe.LK_GEMM and L["ntasks"] > 0 and
                       ((P.gemm[L["task0"]:L["task0"] + L["ntasks"]]["flags"] & pe.GF_ATOMIC) != 0).any())
        assert one_tile > 0
    Ad = A.toarray()
    sign, ld0 = np.linalg.slogdet(Ad)
    assert abs(em.logdet() - ld0) < 1e-11 * abs(ld0)
    B = rng.normal(size=(n, 6))
    assert np.abs(Ad @ em.solve(B, mode=15) - B).max() < 1e-10 * np.abs(B).max()
    Zq = em.selinv()
    Zd = np.linalg.inv(Ad)
    full = pat.to_csc(Zq).toarray()
    mask = pat.to_csc(np.ones(pat.nslots * n)).toarray() != 0
    assert np.abs(full[mask] - Zd[mask]).max() < 1e-10 * np.abs(Zd).max()


def test_emulated_potrf_overlap_schedule(monkeypatch):
    """Opt-in factor schedule with the left-looking update split into its diagonal tile (main lane) and the rows below
    (side lane, LK_SYNC fork / join records): the launch list stays a valid serial order and gives the same factor."""
    M, N, T, bc = 24, 22, 9, 3
    pat = Pattern(M, N, T, bc)
    rng = np.random.default_rng(5)
    plan0 = _lib.PlanHandle(M, N, T, bc)
    n = plan0.n
    W = pat.to_csc(rng.normal(size=pat.nslots * n))
    A = (W + W.T) * 0.5
    A = sparse.csc_matrix(A + sparse.diags(np.abs(A).sum(axis=1).A1 + 1.0))
    em0 = pe.Emulator(plan0)
    assert em0.factorize(pat.from_sparse(A)) == 0
    monkeypatch.setenv("SPDE_POTRF_OVERLAP", "1")
    plan1 = _lib.PlanHandle(M, N, T, bc)
    L1 = pe.Program(plan1, 0).launches
    assert ((L1["kind"] == pe.LK_SYNC) & (L1["variant"] == 1)).sum() > 0 and (L1["lane"] == 1).sum() > 0
    em1 = pe.Emulator(plan1)
    assert em1.factorize(pat.from_sparse(A)) == 0
    assert np.abs(em1.sp[0] - em0.sp[0]).max() <= 1e-13 * np.abs(em0.sp[0]).max()


def test_emulated_selinv_split_k(monkeypatch):
    """The split-K variant of the skinny Takahashi product (normally only for fronts >= 2048 rows)."""
    monkeypatch.setenv("SPDE_SPLITK_MIN", "64")
    monkeypatch.setenv("SPDE_SELINV_OUTER", "0")          # the 64-column recursion (what the streamed schedules use)
    M, N, T, bc = 20, 18, 6, 3
    plan = _lib.PlanHandle(M, N, T, bc)
    pat = Pattern(M, N, T, bc)
    n = plan.n
    rng = np.random.default_rng(2)
    W = pat.to_csc(rng.normal(size=pat.nslots * n))
    A = (W + W.T) * 0.5
    A = sparse.csc_matrix(A + sparse.diags(np.abs(A).sum(axis=1).A1 + 1.0))
    em = pe.Emulator(plan)
    assert em.factorize(pat.from_sparse(A)) == 0
    prog = pe.Program(plan, 3)
    assert ((prog.gemm["flags"] & pe.GF_ATOMIC) != 0).sum() > (prog.gemm["flags"] & pe.GF_MIRROR != 0).sum() // 4
    Zq = em.selinv()
    Zd = np.linalg.inv(A.toarray())
    full = pat.to_csc(Zq).toarray()
    mask = pat.to_csc(np.ones(pat.nslots * n)).toarray() != 0
    assert np.abs(full[mask] - Zd[mask]).max() < 1e-10 * np.abs(Zd).max()


@pytest.mark.parametrize("outer", ["1", "2"])
def test_emulated_blocked_multi_rhs_solves(outer, monkeypatch):
    """Solves with more than four right-hand sides use the two-level blocked schedule (outer blocks of pivot columns);
    SPDE_SOLVE_OUTER shrinks the outer block so that small meshes exercise it."""
    from spdepy_b200 import _lib
    monkeypatch.setenv("SPDE_SOLVE_OUTER", outer)
    monkeypatch.setenv("SPDE_FACTOR_OUTER", outer)      # also the two-lane look-ahead factor schedule (near / far / U pieces)
    monkeypatch.setenv("SPDE_LOOKAHEAD", "1")
    M, N, T = 18, 16, 5
    plan = _lib.PlanHandle(M, N, T, 3)
    first, rowptr, rows, parent = plan.supernodes()
    assert np.diff(first).max() > 64 * int(outer), "mesh too small to reach the outer-block path"
    n = plan.n
    from spdepy_b200.pattern import Pattern
    pat = Pattern(M, N, T, 3)
    rng = np.random.default_rng(1)
    # a symmetric, strictly diagonally dominant matrix on the Q43 pattern
    ones = pat.to_csc(np.ones(43 * n))
    A = ones.multiply(1.0).tocsr()
    A.data[:] = rng.uniform(-1.0, 1.0, size=A.data.size)
    A = (A + A.T) * 0.5
    A = A + sparse.diags(np.asarray(abs(A).sum(axis=1)).ravel() + 1.0)
    flat = pat.from_sparse(A.tocsc())
    em = pe.Emulator(plan)
    launches = pe.Program(plan, 0).launches
    assert (launches["kind"] == pe.LK_SYNC).sum() > 0
    if outer == "1":
        assert (launches["lane"] == 1).sum() > 0
    assert em.factorize(flat, None, 0.0) == 0
    Ad = A.toarray()
    sign, ld = np.linalg.slogdet(Ad)
    assert abs(em.logdet() - ld) < 1e-10 * abs(ld)
    B = rng.normal(size=(n, 6))
    X = em.solve(B, mode=15)
    assert np.abs(X - np.linalg.solve(Ad, B)).max() < 1e-9 * np.abs(X).max()
    perm = plan.perm.astype(np.int64)
    Lref = np.linalg.cholesky(Ad[np.ix_(perm, perm)])
    Zs = em.solve(B, mode=10)
    ref = np.empty_like(B)
    ref[perm] = np.linalg.solve(Lref.T, B)
    assert np.abs(Zs - ref).max() < 1e-9 * np.abs(ref).max()
