"""CPU validation of the streamed evaluator (csrc/ooc.cu): segmentation of the supernodal tree, the static pool
plan and the per-segment schedules, executed by the NumPy interpreter (oracle/plan_emulator.py, memory above the
stack poisoned with NaN after every segment) and compared with dense LAPACK.  No CUDA call is made here."""
import numpy as np
import pytest
from scipy import sparse

import plan_emulator as pe
from helpers import load_golden, make_oracle
from spdepy_b200 import _lib
from spdepy_b200.pattern import Pattern

# (golden case, top_bytes): every supernode on top / mixed top + recomputed subtrees / one in-core segment
CASES = [("ad_ani_bc3_q0", 0), ("ad_ani_bc3_q0", 60000), ("ad_ani_bc3_q0", 10 ** 12), ("ad_ha_bc1_q0", 40000),
         ("vavd_ani_bc2", 30000), ("wm_ani_bc1_ext", 8000), ("sep_ani_bc3", 50000)]


def _setup(name):
    d = load_golden(name)
    mod = make_oracle(d)
    mod.setQ(d["par"])
    M, N = mod.grid.shape[0], mod.grid.shape[1]
    timed = d["T"] is not None
    T = mod.grid.T if timed else 1
    pat_id = 1 if d["spde"] == "seperable-spatial-temporal" else 0
    plan = _lib.PlanHandle(M, N, T, d["bc"], pat_id)
    pat = Pattern(M, N, T, d["bc"], pat_id)
    return mod, plan, pat


@pytest.mark.parametrize("name,thr", CASES)
def test_streamed_factor_solve_selinv(name, thr):
    mod, plan, pat = _setup(name)
    n = plan.n
    flat = pat.from_sparse(mod.Q)
    rng = np.random.default_rng(0)
    cnt = np.zeros(n)
    cnt[rng.choice(n, n // 3, replace=False)] = 1.0
    tau = 7.0
    A = sparse.tril(mod.Q).toarray()
    A = A + np.tril(A, -1).T + np.diag(cnt * tau)
    ooc = _lib.OocHandle(plan, thr)
    st = ooc.stats()
    if thr == 0:
        assert st["top_segments"] == st["segments"] == plan.info(1) and st["recompute_flops"] == 0
    if thr >= 10 ** 12:
        assert st["segments"] == 1 and st["host_bytes"] == 0
    assert st["pool_bytes"] >= max(st["peak_forward_bytes"], st["peak_backward_bytes"])
    em = pe.OocEmulator(plan, ooc)
    B = rng.normal(size=(n, 3))
    ld, X, Zq = em.evaluate(flat, cnt, tau, B, 15, True)
    assert em.status == 0
    sign, ld0 = np.linalg.slogdet(A)
    assert abs(ld - ld0) < 1e-11 * abs(ld0)
    X0 = np.linalg.solve(A, B)
    assert np.abs(X - X0).max() < 1e-10 * np.abs(X0).max()
    Zd = np.linalg.inv(A)
    full = pat.to_csc(Zq).toarray()
    mask = pat.to_csc(np.ones(pat.nslots * n)).toarray() != 0
    assert np.abs(full[mask] - Zd[mask]).max() < 1e-10 * np.abs(Zd).max()


def test_streamed_forward_only():
    """want_backward = 0: smaller pool, no host memory; log-determinant and L^-1 P b (the quadratic form
    b^T A^-1 b = |L^-1 P b|^2 that logLike(grad=False) uses)."""
    mod, plan, pat = _setup("ad_ani_bc3_q0")
    n = plan.n
    flat = pat.from_sparse(mod.Q)
    A = sparse.tril(mod.Q).toarray()
    A = A + np.tril(A, -1).T
    both = _lib.OocHandle(plan, 30000, True).stats()
    ooc = _lib.OocHandle(plan, 30000, False)
    st = ooc.stats()
    assert st["host_bytes"] == 0 and st["pool_bytes"] == st["peak_forward_bytes"] <= both["pool_bytes"]
    em = pe.OocEmulator(plan, ooc)
    b = np.random.default_rng(3).normal(size=(n, 2))
    ld, y, _ = em.evaluate(flat, None, 0.0, b, 1 | 4, False)
    sign, ld0 = np.linalg.slogdet(A)
    assert abs(ld - ld0) < 1e-11 * abs(ld0)
    q = np.einsum("ij,ij->", b, np.linalg.solve(A, b))
    assert abs((y * y).sum() - q) < 1e-11 * abs(q)


@pytest.mark.parametrize("shape", [(24, 22, 9, 3, 150000), (21, 20, 6, 2, 100000)])
def test_streamed_large_fronts(shape):
    """Multi-block supernodes on top (outer right-looking blocks, split-K Takahashi) with recomputed subtrees."""
    M, N, T, bc, thr = shape
    plan = _lib.PlanHandle(M, N, T, bc)
    pat = Pattern(M, N, T, bc)
    n = plan.n
    rng = np.random.default_rng(1)
    W = pat.to_csc(rng.normal(size=pat.nslots * n))
    A = (W + W.T) * 0.5
    A = sparse.csc_matrix(A + sparse.diags(np.abs(A).sum(axis=1).A1 + 1.0))
    flat = pat.from_sparse(A)
    ooc = _lib.OocHandle(plan, thr)
    st = ooc.stats()
    assert 0 < st["top_segments"] < st["segments"]
    em = pe.OocEmulator(plan, ooc)
    B = rng.normal(size=(n, 1))
    ld, X, Zq = em.evaluate(flat, None, 0.0, B, 15, True)
    Ad = A.toarray()
    sign, ld0 = np.linalg.slogdet(Ad)
    assert abs(ld - ld0) < 1e-11 * abs(ld0)
    assert np.abs(Ad @ X - B).max() < 1e-10 * np.abs(B).max()
    Zd = np.linalg.inv(Ad)
    full = pat.to_csc(Zq).toarray()
    mask = pat.to_csc(np.ones(pat.nslots * n)).toarray() != 0
    assert np.abs(full[mask] - Zd[mask]).max() < 1e-10 * np.abs(Zd).max()


def test_memory_plan_of_the_headline_mesh_is_monotone():
    """The pool never grows when more supernodes are handled front by front, and the host pool never shrinks."""
    plan = _lib.PlanHandle(40, 40, 16, 3)
    prev = None
    for thr in (10 ** 12, 4 * 10 ** 6, 10 ** 6, 250000, 0):
        st = _lib.OocHandle(plan, thr, True, False).stats()
        if prev is not None:
            assert st["pool_bytes"] <= prev["pool_bytes"] and st["host_bytes"] >= prev["host_bytes"]
            assert st["recompute_flops"] <= prev["recompute_flops"] or prev["segments"] == 1
        prev = st


def _dense_case(M, N, T, bc, seed=1):
    plan = _lib.PlanHandle(M, N, T, bc)
    pat = Pattern(M, N, T, bc)
    n = plan.n
    rng = np.random.default_rng(seed)
    W = pat.to_csc(rng.normal(size=pat.nslots * n))
    A = (W + W.T) * 0.5
    A = sparse.csc_matrix(A + sparse.diags(np.abs(A).sum(axis=1).A1 + 1.0))
    return plan, pat, A, rng


@pytest.mark.parametrize("outer,overlap,selinv,selblocks", [(1, 1, True, 1), (1, 1, False, 1), (2, 1, True, 2), (1, 0, True, 1),
                                                            (1, 1, True, 3), (2, 1, True, 1), (1, 1, True, 0)])
def test_streamed_panel_slices(monkeypatch, outer, overlap, selinv, selblocks):
    """Overlapped panel traffic of the spilled top segments (LK_COPY records): with one 64-column block per outer block
    the fronts on top are parked in several slices during their factorisation and fetched back slice by slice, last
    slice first, by the wait records inside the Takahashi schedule; the interpreter fetches a slice only at its wait
    record and everything else above the stack is NaN, so a block column read too early poisons the result.  Also
    without the selected inverse (the solve waits for all slices) and with the overlap switched off."""
    monkeypatch.setenv("SPDE_FACTOR_OUTER", str(outer))
    monkeypatch.setenv("SPDE_OOC_OVERLAP", str(overlap))
    # outer blocks of the two-level Takahashi recursion: as wide as the slices, wider, narrower; 0 = the 64-column recursion
    if selblocks:
        monkeypatch.setenv("SPDE_SELINV_OUTER_BLOCKS", str(selblocks))
    else:
        monkeypatch.setenv("SPDE_SELINV_OUTER", "0")
    plan, pat, A, rng = _dense_case(24, 22, 9, 3)
    n = plan.n
    ooc = _lib.OocHandle(plan, 150000)
    st = ooc.stats()
    nsl = [max(len(ooc.export(s, 0, 7, "i8")) // 3 - 1, 0) for s in range(st["segments"])]
    if overlap:
        assert max(nsl) >= 4 // outer and sum(1 for v in nsl if v) == st["top_segments"] - 1   # all but the kept root
        for s in range(st["segments"]):
            if nsl[s]:
                park = ooc.export(s, 0, 0, pe.LAUNCH)
                wait = ooc.export(s, 3, 0, pe.LAUNCH)
                assert (park["kind"] == pe.LK_COPY).sum() == nsl[s] + 1          # slices + inverse diagonal blocks
                nwait = ((wait["kind"] == pe.LK_COPY) & (wait["variant"] == 1)).sum()
                if selblocks in (0, outer):
                    assert nwait == nsl[s]                  # one wait record per slice
                else:
                    assert 1 <= nwait                        # one per outer block of the recursion (several may name one slice)
    else:
        assert max(nsl) == 0
    em = pe.OocEmulator(plan, ooc)
    B = rng.normal(size=(n, 2))
    ld, X, Zq = em.evaluate(pat.from_sparse(A), None, 0.0, B, 15, selinv)
    Ad = A.toarray()
    sign, ld0 = np.linalg.slogdet(Ad)
    assert abs(ld - ld0) < 1e-11 * abs(ld0)
    assert np.abs(Ad @ X - B).max() < 1e-10 * np.abs(B).max()
    if selinv:
        Zd = np.linalg.inv(Ad)
        full = pat.to_csc(Zq).toarray()
        mask = pat.to_csc(np.ones(pat.nslots * n)).toarray() != 0
        assert np.abs(full[mask] - Zd[mask]).max() < 1e-10 * np.abs(Zd).max()


def test_streamed_forward_only_ignores_park_records(monkeypatch):
    """A plan built with a backward pass, run forward only: the park records in the factor schedules are skipped."""
    monkeypatch.setenv("SPDE_FACTOR_OUTER", "1")
    plan, pat, A, rng = _dense_case(21, 20, 6, 2, seed=5)
    ooc = _lib.OocHandle(plan, 100000)
    em = pe.OocEmulator(plan, ooc)
    b = rng.normal(size=(plan.n, 1))
    ld, y, _ = em.evaluate(pat.from_sparse(A), None, 0.0, b, 1 | 4, False)
    assert em.hostbuf is None
    Ad = A.toarray()
    assert abs(ld - np.linalg.slogdet(Ad)[1]) < 1e-11 * abs(ld)
    q = float(b[:, 0] @ np.linalg.solve(Ad, b[:, 0]))
    assert abs((y * y).sum() - q) < 1e-11 * abs(q)
