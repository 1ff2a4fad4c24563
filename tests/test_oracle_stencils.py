"""The C restatement of the stencil generators (oracle/stencil_oracle.c) must be bit-identical to
the reference's own C++ (oracle/_ref/lib_*.so, compiled from /root/reference by oracle/Makefile)."""
import numpy as np
import pytest

import stencils as st

pytestmark = pytest.mark.skipif(not st.ref_available(), reason="oracle/_ref not built")

SHAPES = [(7, 5), (12, 10), (3, 3), (2, 6)]


def _same(a, b):
    for x, y in zip(a, b):
        assert x.dtype == y.dtype and x.shape == y.shape
        if x.dtype == np.float64:
            assert np.array_equal(x.view(np.int64), y.view(np.int64)) or \
                np.array_equal(np.nan_to_num(x, nan=1e300), np.nan_to_num(y, nan=1e300))
        else:
            assert np.array_equal(x, y)


@pytest.mark.parametrize("bc", [1, 3])
@pytest.mark.parametrize("shape", SHAPES)
def test_ah_const(bc, shape):
    M, N = shape
    rng = np.random.default_rng(M * 100 + N + bc)
    for _ in range(3):
        v = rng.normal(size=2)
        H = np.exp(rng.normal()) * np.eye(2) + np.outer(v, v)
        _same(st.oracle_ah_const(M, N, H, 0.3, 0.7, bc), st.ref_ah_const(M, N, H, 0.3, 0.7, bc))


def test_ah_const_periodic_rejected():
    with pytest.raises(ValueError):
        st.oracle_ah_const(5, 5, np.eye(2), 1.0, 1.0, 2)


@pytest.mark.parametrize("bc", [1, 2, 3])
@pytest.mark.parametrize("shape", SHAPES)
def test_ah_face(bc, shape):
    M, N = shape
    rng = np.random.default_rng(M * 100 + N + bc)
    H = rng.normal(size=(M * N, 4, 2, 2))
    H0 = H.copy()
    _same(st.oracle_ah_face(M, N, H, 0.3, 0.7, bc), st.ref_ah_face(M, N, H, 0.3, 0.7, bc))
    assert np.array_equal(H, H0)


@pytest.mark.parametrize("bc", [1, 2, 3])
@pytest.mark.parametrize("diff", [1, 2, 3])
@pytest.mark.parametrize("shape", SHAPES)
def test_aw_const(bc, diff, shape):
    M, N = shape
    for G in ([0.4, -1.3], [-2.0, 0.5], [1.0, 1.0]):
        _same(st.oracle_aw_const(M, N, G, 0.3, 0.7, diff, bc), st.ref_aw_const(M, N, G, 0.3, 0.7, diff, bc))


@pytest.mark.parametrize("bc", [1, 2, 3])
@pytest.mark.parametrize("diff", [1, 2, 3])
@pytest.mark.parametrize("shape", SHAPES)
def test_aw_face(bc, diff, shape):
    M, N = shape
    rng = np.random.default_rng(M * 100 + N + bc + diff)
    G = rng.normal(size=(M * N, 4))
    G[rng.random(size=G.shape) < 0.1] = 0.0          # exact zeros -> 0/0 = NaN in diff modes
    dG = rng.normal(size=(M * N, 4))
    _same(st.oracle_aw_face(M, N, G, dG, 0.3, 0.7, diff, bc), st.ref_aw_face(M, N, G, dG, 0.3, 0.7, diff, bc))
