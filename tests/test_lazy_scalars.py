"""Lazy scalars of the device-resident likelihood epilogue (spdepy_b200/engine.py: Lazy): the arithmetic of
``advection_diffusion2D.py:198-223`` written on pool slots must evaluate, after ONE fetch, to exactly what NumPy floats
give -- including reflected operators with NumPy scalars, which must not turn a Lazy into an object array."""
import numpy as np

from spdepy_b200.engine import Lazy


class FakePool:
    def __init__(self, values):
        self.values = np.asarray(values, dtype=np.float64)
        self.fetches = 0

    def fetch(self):
        self.fetches += 1
        return self.values

    def scalar(self, i):
        return Lazy(self, lambda h: h[i])


def test_lazy_arithmetic_matches_floats():
    h = np.array([3.25, -1.5, 7.0, 0.125])
    pool = FakePool(h)
    a, b, c, d = (pool.scalar(i) for i in range(4))
    r, nobs, tau = 20, 5000, np.float64(1000.0)
    like = 1 / 2 * a * r + nobs * r * np.log(tau) / 2 - 1 / 2 * b * r - 1 / 2 * c - tau / 2 * d
    ref = 1 / 2 * h[0] * r + nobs * r * np.log(tau) / 2 - 1 / 2 * h[1] * r - 1 / 2 * h[2] - tau / 2 * h[3]
    assert pool.fetches == 0                      # nothing is read before the first float()
    assert float(like) == ref
    g = nobs * r / 2 - 1 / 2 * (a * tau) * r - tau / 2 * d
    assert float(g) == nobs * r / 2 - 1 / 2 * (h[0] * tau) * r - tau / 2 * h[3]
    assert float(-(a - b) + 2.0) == -(h[0] - h[1]) + 2.0
    assert float((a + b) / c) == (h[0] + h[1]) / h[2]
    assert float(5 - a) == 5 - h[0] and float(np.float64(5) - a) == 5 - h[0]
    assert float(np.float64(2.0) * a) == 2.0 * h[0] and isinstance(np.float64(2.0) * a, Lazy)
    assert float(0.0 + a) == h[0]
