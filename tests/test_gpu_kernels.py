"""GPU parity tests of the individual kernels, through the C ABI: stencils and the assembled Q
against the oracle (bit-exact / 1e-13), the grouped FP64 tensor-core GEMM against a plain torch
fp64 matmul (the one floating-point kernel with a torch reference)."""
import ctypes

import numpy as np
import pytest
import torch

import stencils as st
from helpers import canon, golden_names, load_golden, make_grids, make_oracle, relerr, spec_key

pytestmark = pytest.mark.gpu


def _eng(M, N, T, bc):
    from spdepy_b200.engine import Engine
    return Engine.get(M, N, T, bc)


def _dev(a):
    from spdepy_b200.engine import to_dev
    return to_dev(a)


def _csc_of_slots(model, a9):
    return model._stencil_to_csc(a9)


@pytest.mark.parametrize("bc", [1, 3])
@pytest.mark.parametrize("shape", [(7, 5), (12, 10), (33, 17)])
def test_ah_const_bit_exact(bc, shape):
    M, N = shape
    eng = _eng(M, N, 1, bc)
    rng = np.random.default_rng(M + bc)
    v = rng.normal(size=2)
    H = np.exp(rng.normal()) * np.eye(2) + np.outer(v, v)
    got = eng.ah_stencil(0.3, 0.7, _dev(H), False).cpu().numpy().reshape(9, M * N)
    ref = st.to_csc(st.oracle_ah_const(M, N, H, 0.3, 0.7, bc), M * N).toarray()
    _check_slots(got, ref, M, N, bc)


@pytest.mark.parametrize("bc", [1, 2, 3])
@pytest.mark.parametrize("shape", [(7, 5), (12, 10), (33, 17)])
def test_ah_face_bit_exact(bc, shape):
    M, N = shape
    eng = _eng(M, N, 1, bc)
    H = np.random.default_rng(M + bc).normal(size=(M * N, 4, 2, 2))
    got = eng.ah_stencil(0.3, 0.7, _dev(H), True).cpu().numpy().reshape(9, M * N)
    ref = st.to_csc(st.oracle_ah_face(M, N, H, 0.3, 0.7, bc), M * N).toarray()
    _check_slots(got, ref, M, N, bc)


@pytest.mark.parametrize("bc", [1, 2, 3])
@pytest.mark.parametrize("diff", [1, 2, 3])
def test_aw_bit_exact(bc, diff):
    M, N = 12, 10
    eng = _eng(M, N, 1, bc)
    rng = np.random.default_rng(bc * 10 + diff)
    G = np.array([0.4, -1.3])
    got = eng.aw_stencil(0.3, 0.7, _dev(G), None, False, diff, False).cpu().numpy().reshape(9, M * N)
    ref = st.to_csc(st.oracle_aw_const(M, N, G, 0.3, 0.7, diff, bc), M * N).toarray()
    _check_slots(got, ref, M, N, bc)
    Gf = rng.normal(size=(M * N, 4))
    Gf[rng.random(size=Gf.shape) < 0.1] = 0.0
    dG = rng.normal(size=(M * N, 4))
    got = eng.aw_stencil(0.3, 0.7, _dev(Gf), _dev(dG), True, diff, True).cpu().numpy().reshape(9, M * N)
    ref = st.to_csc(st.oracle_aw_face(M, N, Gf, dG, 0.3, 0.7, diff, bc), M * N, nan_to_zero=True).toarray()
    _check_slots(got, ref, M, N, bc)


def _check_slots(got, ref, M, N, bc):
    k = np.arange(M * N)
    i, j = k % M, k // M
    dense = np.zeros((M * N, M * N))
    for s in range(9):
        ii, jj = i + (s % 3 - 1), j + (s // 3 - 1)
        if bc == 2:
            ii, jj = ii % M, jj % N
            ok = np.ones(M * N, bool)
        else:
            ok = (ii >= 0) & (ii < M) & (jj >= 0) & (jj < N)
            assert np.all(got[s][~ok] == 0.0)
        dense[k[ok], (jj * M + ii)[ok]] = got[s][ok]
    assert np.array_equal(dense, ref), np.abs(dense - ref).max()


@pytest.mark.parametrize("name", golden_names())
def test_makeQ_against_reference_golden(name):
    """Q from the CUDA assembly vs what the unmodified reference produced: pattern exact after
    eliminate_zeros(), values to 1e-13 relative (BASELINE.json north_star)."""
    import spdepy_b200 as sp
    d = load_golden(name)
    g, g0 = make_grids(d)
    kw = {}
    if g0 is not None:
        m0 = sp.model(grid=g0, spde=d["mod0_spde"], ha=d["ha"], anisotropic=d["ani"], bc=d["bc"], parameters=d["mod0_par"])
        kw["mod0"] = m0
    mod = sp.model(grid=g, spde=d["spde"], ha=d["ha"], anisotropic=d["ani"], bc=d["bc"], **kw)
    if "ww" in d and d["ww"].size:
        mod.mod.ww = d["ww"]
    assert mod.mod.type == d["type"]
    Q = canon(mod.mod.makeQ(d["par"], grad=False)[0])        # (Q, Q_fac, None); the separable class returns (Q, Q_fac)
    ref = d["Q"]
    assert np.array_equal(Q.indptr, ref.indptr) and np.array_equal(Q.indices, ref.indices)
    rel = np.abs(Q.data - ref.data) / np.abs(ref.data)
    assert rel.max() <= 1e-13, rel.max()


GEMM_CASES = [
    # cfg, akmaj, bkmaj, M, N, K, flags
    (0, 0, 0, 300, 260, 100, 1 << 11),
    (0, 0, 0, 257, 129, 37, 0),
    (1, 0, 0, 1000, 64, 64, 1 << 10),
    (2, 0, 0, 70, 50, 23, (1 << 11) | (1 << 9)),
    (2, 0, 1, 66, 64, 129, 1 << 11),
    (0, 0, 1, 256, 256, 64, 0),
    (2, 1, 1, 64, 64, 1000, 1 << 11),
    (1, 1, 0, 130, 60, 77, 0),
    (0, 0, 0, 128, 128, 16, 1 << 10),
    (2, 0, 0, 2, 2, 2, 0),
]


@pytest.mark.parametrize("case", GEMM_CASES)
def test_grouped_gemm_vs_torch(case):
    from spdepy_b200._lib import check, lib
    cfg, ak, bk, M, N, K, flags = case
    dev = torch.device("cuda")
    gen = torch.Generator(device="cuda").manual_seed(1)
    ev = lambda v: v + (v & 1)
    lda = ev(K if ak else M) + 2          # the library keeps every leading dimension even (16-byte cp.async)
    ldb = ev(K if bk else N) + 4
    ldc = ev(M) + 2
    A = torch.randn((M if ak else K), lda, dtype=torch.float64, device=dev, generator=gen)   # column-major storage: [col][ld]
    B = torch.randn((N if bk else K), ldb, dtype=torch.float64, device=dev, generator=gen)
    C = torch.randn(N, ldc, dtype=torch.float64, device=dev, generator=gen)
    Am = (A[:, :K]) if ak else A[:, :M].T              # logical M x K
    Bm = (B[:, :K].T) if bk else B[:, :N]              # logical K x N
    C0 = C.clone()
    ms = ctypes.c_float()
    check(lib.spde_gemm_single(cfg, ak, bk, flags, M, N, K, A.data_ptr(), lda, B.data_ptr(), ldb, C.data_ptr(), ldc, 0,
                               ctypes.byref(ms), None))
    torch.cuda.synchronize()
    prod = Am @ Bm                                     # torch fp64 reference
    if flags & (1 << 11):
        prod = -prod
    ref = C0.clone()
    want = prod.T if (flags & (1 << 10)) else C0[:, :M] + prod.T
    if flags & (1 << 9):
        mask = (torch.arange(M, device=dev)[None, :] >= torch.arange(N, device=dev)[:, None])
        ref[:, :M] = torch.where(mask, want, C0[:, :M])
    else:
        ref[:, :M] = want
    scale = float(prod.abs().max()) + 1.0
    assert float((C - ref).abs().max()) <= 1e-12 * scale * max(1, K // 16)
