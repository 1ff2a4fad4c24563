"""GPU parity tests of the streamed (depth-first, statically planned) evaluator, through the C ABI (spde_ooc_*):
against dense LAPACK on small meshes, against the in-core path on a mid-size mesh, and logLike / exact gradient
against the oracle with the streamed path forced."""
import numpy as np
import pytest
import torch
from scipy import sparse

from helpers import load_golden, make_grids, make_oracle, relerr

pytestmark = pytest.mark.gpu


def _eng(M, N, T, bc):
    from spdepy_b200.engine import Engine
    return Engine.get(M, N, T, bc)


def _spd_on_pattern(eng, seed):
    pat = eng.pattern
    n = eng.n
    rng = np.random.default_rng(seed)
    W = pat.to_csc(rng.normal(size=pat.nslots * n))
    A = (W + W.T) * 0.5
    A = sparse.csc_matrix(A + sparse.diags(np.abs(A).sum(axis=1).A1 + 1.0))
    return A, pat.from_sparse(A)


@pytest.mark.parametrize("shape", [(24, 22, 9, 3), (21, 20, 6, 2), (40, 37, 1, 1), (16, 14, 6, 1)])
@pytest.mark.parametrize("thr", [0, 120000, 10 ** 12])
@pytest.mark.parametrize("k", [1, 3, 40])
def test_streamed_vs_dense(shape, thr, k, monkeypatch):
    from spdepy_b200 import _lib
    from spdepy_b200.engine import to_dev
    M, N, T, bc = shape
    eng = _eng(M, N, T, bc)
    n = eng.n
    A, flat = _spd_on_pattern(eng, M + k)
    rng = np.random.default_rng(k)
    cnt = np.zeros(n)
    cnt[rng.choice(n, n // 4, replace=False)] = 1.0
    tau = 3.0
    Ad = A.toarray() + np.diag(cnt * tau)
    ooc = _lib.OocHandle(eng.plan, thr, True, True)
    B = rng.normal(size=(n, k))
    X = to_dev(B.copy())
    Z = torch.empty(eng.nslots * n, dtype=torch.float64, device="cuda")
    Qd, cd = to_dev(flat), to_dev(cnt)          # keep the device buffers alive across the call
    ld = ooc.run(Qd.data_ptr(), cd.data_ptr(), tau, X.data_ptr(), k, 15, Z.data_ptr(),
                 torch.cuda.current_stream().cuda_stream)
    sign, ld0 = np.linalg.slogdet(Ad)
    assert abs(ld - ld0) <= 1e-11 * abs(ld0)
    assert relerr(X.cpu().numpy(), np.linalg.solve(Ad, B)) <= 1e-9
    Zd = np.linalg.inv(Ad)
    full = eng.pattern.to_csc(Z.cpu().numpy()).toarray()
    mask = eng.pattern.to_csc(np.ones(eng.nslots * n)).toarray() != 0
    assert np.abs(full[mask] - Zd[mask]).max() <= 1e-9 * np.abs(Zd).max()
    # a second run on the same handle (pool reuse, stale memory from the previous pass)
    X2 = to_dev(B.copy())
    ld2 = ooc.run(Qd.data_ptr(), cd.data_ptr(), tau, X2.data_ptr(), k, 15, Z.data_ptr(),
                  torch.cuda.current_stream().cuda_stream)
    assert abs(ld2 - ld) <= 1e-13 * abs(ld) and relerr(X2.cpu().numpy(), X.cpu().numpy()) <= 1e-12


def test_streamed_forward_only_and_not_spd():
    from spdepy_b200 import _lib
    from spdepy_b200.engine import to_dev
    eng = _eng(24, 22, 9, 3)
    n = eng.n
    A, flat = _spd_on_pattern(eng, 5)
    ooc = _lib.OocHandle(eng.plan, 100000, False, True)
    assert ooc.stats()["host_bytes"] == 0
    b = np.random.default_rng(2).normal(size=(n, 2))
    y = to_dev(b.copy())
    st = torch.cuda.current_stream().cuda_stream
    Qd = to_dev(flat)
    ld = ooc.run(Qd.data_ptr(), None, 0.0, y.data_ptr(), 2, 1 | 4, None, st)
    Ad = A.toarray()
    assert abs(ld - np.linalg.slogdet(Ad)[1]) <= 1e-11 * abs(ld)
    q = np.einsum("ij,ij->", b, np.linalg.solve(Ad, b))
    assert abs(float((y * y).sum()) - q) <= 1e-10 * abs(q)
    with pytest.raises(ValueError):      # a backward pass on a forward-only plan
        ooc.run(Qd.data_ptr(), None, 0.0, y.data_ptr(), 2, 15, None, st)
    bad = flat.copy()
    bad[(eng.nslots // 2) * n + 7] = -1.0
    Qbad = to_dev(bad)
    with pytest.raises(_lib.NotPositiveDefiniteError):
        ooc.run(Qbad.data_ptr(), None, 0.0, None, 0, 0, None, st)


def test_streamed_vs_incore_midsize():
    """64x64x25 (n = 1.0e5, 0.8 GB of factor): the streamed pass with a 40 MB threshold against the in-core stores."""
    from spdepy_b200 import _lib
    eng = _eng(64, 64, 25, 3)
    n = eng.n
    A, flat = _spd_on_pattern(eng, 11)
    Q = torch.as_tensor(flat, device="cuda")
    gen = torch.Generator(device="cuda").manual_seed(3)
    B = torch.randn(n, 2, dtype=torch.float64, device="cuda", generator=gen)
    eng.factorize(1, Q)
    ld0 = eng.logdet(1)
    X0 = eng.solve(1, B.clone())
    Z0 = eng.selinv(1)
    ooc = _lib.OocHandle(eng.plan, 40 * 10 ** 6, True, True)
    st = ooc.stats()
    assert 0 < st["top_segments"] < st["segments"] and st["pool_bytes"] < eng.incore_bytes()
    X = B.clone()
    Z = torch.empty_like(Z0)
    ld = ooc.run(Q.data_ptr(), None, 0.0, X.data_ptr(), 2, 15, Z.data_ptr(), torch.cuda.current_stream().cuda_stream)
    assert abs(ld - ld0) <= 1e-12 * abs(ld0)
    assert float((X - X0).abs().max()) <= 1e-11 * float(X0.abs().max())
    assert float((Z - Z0).abs().max()) <= 1e-11 * float(Z0.abs().max())
    R = eng.q_apply(Q, X) - B                      # residual of the streamed solve
    assert float(R.abs().max()) <= 1e-10 * float(B.abs().max())


@pytest.mark.parametrize("name", ["ad_ani_bc3_q0", "ad_ha_bc1_q0", "vavd_ani_bc1_ext_q0", "avd_ani_bc3"])
def test_streamed_loglike_vs_oracle(name, monkeypatch):
    """logLike with the streamed path forced (collapsed prior + depth-first posterior): value and exact gradient
    against the oracle's dense-inverse formula, and the gradient-free value (forward pass only) against it."""
    from test_gpu_loglike import _build
    d = load_golden(name)
    m = _build(d).mod
    m.initFit(d["data"], idx=d["idx"], fitQ0=d["fitQ0"])
    orc = make_oracle(d)
    orc.initFit(d["data"], idx=d["idx"])
    like_o, jac_o = orc.logLike_exact(d["par"])
    monkeypatch.setenv("SPDE_OOC_TOP_BYTES", "30000")
    eng = m.engine
    eng.streamed = True
    m.check_selinv = True
    try:
        like, jac = m.logLike(d["par"], grad=True, exact_grad=True)
        assert eng._ooc is not None and eng._ooc.stats()["top_segments"] > 0
        assert abs(m.last["selinv_trace_over_n"] - 1.0) <= 1e-10      # tr(Q_c Z) = n, the full-size checksum of bench c4
        like_only = m.logLike(d["par"], grad=False)
    finally:
        eng.streamed = None
        eng._ooc = None
    assert abs(like - like_o) <= 1e-9 * abs(like_o)
    assert np.abs(jac - jac_o).max() <= 1e-9 * np.abs(jac_o).max()
    assert abs(like_only - like_o) <= 1e-9 * abs(like_o)


@pytest.mark.parametrize("outer", [1, 2])
def test_streamed_panel_slices_overlapped(outer, monkeypatch):
    """Panel traffic of the spilled top segments on the copy stream (LK_COPY records): one 64-column block per outer
    block makes the fronts on top travel in several slices.  Against dense LAPACK; solve without the selected inverse
    (the solve waits for the last slice); a profiled run (everything on one stream); and to rounding of the atomics against the same
    plan with the overlap switched off (SPDE_OOC_OVERLAP=0: whole-panel copies on the compute stream)."""
    from spdepy_b200 import _lib
    from spdepy_b200.engine import to_dev
    eng = _eng(24, 22, 9, 3)
    n = eng.n
    A, flat = _spd_on_pattern(eng, 21)
    Ad = A.toarray()
    monkeypatch.setenv("SPDE_FACTOR_OUTER", str(outer))
    ooc = _lib.OocHandle(eng.plan, 150000, True, True)
    nsl = [max(len(ooc.export(s, 0, 7, "i8")) // 3 - 1, 0) for s in range(ooc.stats()["segments"])]
    assert max(nsl) >= 4 // outer
    monkeypatch.setenv("SPDE_OOC_OVERLAP", "0")
    plain = _lib.OocHandle(eng.plan, 150000, True, True)
    assert max(len(plain.export(s, 0, 7, "i8")) for s in range(plain.stats()["segments"])) == 0
    B = np.random.default_rng(4).normal(size=(n, 3))
    Qd = to_dev(flat)
    st = torch.cuda.current_stream().cuda_stream
    out = {}
    for name, h in (("overlap", ooc), ("plain", plain)):
        X = to_dev(B.copy())
        Z = torch.empty(eng.nslots * n, dtype=torch.float64, device="cuda")
        ld = h.run(Qd.data_ptr(), None, 0.0, X.data_ptr(), 3, 15, Z.data_ptr(), st)
        out[name] = (ld, X.cpu().numpy(), Z.cpu().numpy())
    ld, X, Z = out["overlap"]
    assert abs(ld - np.linalg.slogdet(Ad)[1]) <= 1e-11 * abs(ld)
    assert relerr(X, np.linalg.solve(Ad, B)) <= 1e-9
    Zd = np.linalg.inv(Ad)
    full = eng.pattern.to_csc(Z).toarray()
    mask = eng.pattern.to_csc(np.ones(eng.nslots * n)).toarray() != 0
    assert np.abs(full[mask] - Zd[mask]).max() <= 1e-9 * np.abs(Zd).max()
    assert abs(ld - out["plain"][0]) <= 1e-13 * abs(ld) and relerr(X, out["plain"][1]) <= 1e-12 and relerr(Z, out["plain"][2]) <= 1e-12
    # solve only (no selected inverse), second run on the same handle
    X2 = to_dev(B.copy())
    ld2 = ooc.run(Qd.data_ptr(), None, 0.0, X2.data_ptr(), 3, 15, None, st)
    assert abs(ld2 - ld) <= 1e-13 * abs(ld) and relerr(X2.cpu().numpy(), X) <= 1e-12
    # profiled run: copies stay on the compute stream
    eng.plan.profile(True)
    try:
        X3 = to_dev(B.copy())
        Z3 = torch.empty(eng.nslots * n, dtype=torch.float64, device="cuda")
        ld3 = ooc.run(Qd.data_ptr(), None, 0.0, X3.data_ptr(), 3, 15, Z3.data_ptr(), st)
    finally:
        eng.plan.profile(False)
    assert abs(ld3 - ld) <= 1e-13 * abs(ld) and relerr(X3.cpu().numpy(), X) <= 1e-12 and relerr(Z3.cpu().numpy(), Z) <= 1e-12
