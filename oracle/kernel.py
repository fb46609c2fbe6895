"""SPH smoothing kernels (oracle; test infrastructure only).

Restates jax_sph/kernel.py: QuinticKernel :51-72, WendlandC2Kernel :75-103 (the
two on the hot path) plus Cubic :26-48, WendlandC4 :106-134, WendlandC6
:137-165, Gaussian :168-182, SuperGaussian :185-201 so that the reference's tests/test_kernel.py can be
replayed against the oracle.  ``grad_w`` is the analytic derivative that
``jax.grad`` (kernel.py:20-23) produces; integer powers follow
``lax.integer_pow`` (repeated squaring).
"""

import numpy as np


def _ipow(x, n):
    """x**n as XLA's integer_pow does it (square-and-multiply, LSB first)."""
    acc = None
    base = x
    while n > 0:
        if n & 1:
            acc = base if acc is None else acc * base
        n >>= 1
        if n:
            base = base * base
    return acc


class _Base:
    def __init__(self, h, dim, dtype):
        self.dtype = np.dtype(dtype)
        self.h = float(h)
        self.dim = dim
        self._one_over_h = 1.0 / h  # python double, kernel.py:55
        self.cutoff = self._normalized_cutoff * h

    def _c(self, v):
        return self.dtype.type(v)


class QuinticKernel(_Base):
    _normalized_cutoff = 3.0

    def __init__(self, h, dim=3, dtype=np.float64):
        super().__init__(h, dim, dtype)
        ooh = self._one_over_h
        self._sigma = {
            1: 1.0 / 120.0 * ooh,
            2: 7.0 / 478.0 / np.pi * ooh**2,
            3: 3.0 / 359.0 / np.pi * ooh**3,
        }[dim]

    def _q(self, r):
        c = self._c
        q = r * c(self._one_over_h)
        q1 = np.maximum(c(0.0), c(1.0) - q)
        q2 = np.maximum(c(0.0), c(2.0) - q)
        q3 = np.maximum(c(0.0), c(3.0) - q)
        return q1, q2, q3

    def w(self, r):
        c = self._c
        q1, q2, q3 = self._q(np.asarray(r, dtype=self.dtype))
        return c(self._sigma) * (_ipow(q3, 5) - c(6.0) * _ipow(q2, 5) + c(15.0) * _ipow(q1, 5))

    def grad_w(self, r):
        c = self._c
        q1, q2, q3 = self._q(np.asarray(r, dtype=self.dtype))
        # d/dr max(0, k - r/h)^5 = -5/h * max(0, k - r/h)^4
        poly = c(-5.0) * _ipow(q3, 4) + c(30.0) * _ipow(q2, 4) - c(75.0) * _ipow(q1, 4)
        return c(self._sigma) * c(self._one_over_h) * poly


class WendlandC2Kernel(_Base):
    _normalized_cutoff = 2.0

    def __init__(self, h, dim=3, dtype=np.float64):
        super().__init__(h, dim, dtype)
        ooh = self._one_over_h
        self._sigma = {
            1: 5.0 / 8.0 * ooh,
            2: 7.0 / 4.0 / np.pi * ooh**2,
            3: 21.0 / 16.0 / np.pi * ooh**3,
        }[dim]

    def w(self, r):
        c = self._c
        q = np.asarray(r, dtype=self.dtype) * c(self._one_over_h)
        q1 = np.maximum(c(0.0), c(1.0) - c(0.5) * q)
        if self.dim == 1:
            return c(self._sigma) * (_ipow(q1, 3) * (c(1.5) * q + c(1.0)))
        return c(self._sigma) * (_ipow(q1, 4) * (c(2.0) * q + c(1.0)))

    def grad_w(self, r):
        c = self._c
        q = np.asarray(r, dtype=self.dtype) * c(self._one_over_h)
        q1 = np.maximum(c(0.0), c(1.0) - c(0.5) * q)
        ooh = c(self._one_over_h)
        if self.dim == 1:
            # d/dq [q1^3 (1.5 q + 1)] = -1.5 q1^2 (1.5q+1) + 1.5 q1^3 = -3 q q1^2
            return c(self._sigma) * ooh * (c(-3.0) * q * _ipow(q1, 2))
        # d/dq [q1^4 (2q+1)] = -2 q1^3 (2q+1) + 2 q1^4 = -5 q q1^3
        return c(self._sigma) * ooh * (c(-5.0) * q * _ipow(q1, 3))


class CubicKernel(_Base):
    _normalized_cutoff = 2.0

    def __init__(self, h, dim=3, dtype=np.float64):
        super().__init__(h, dim, dtype)
        ooh = self._one_over_h
        self._sigma = {1: 2.0 / 3.0 * ooh, 2: 10.0 / 7.0 / np.pi * ooh**2, 3: 1.0 / np.pi * ooh**3}[
            dim
        ]

    def w(self, r):
        q = np.asarray(r, dtype=self.dtype) * self._c(self._one_over_h)
        c1 = (1 - q >= 0).astype(self.dtype)
        c2 = ((2 - q < 1) & (2 - q >= 0)).astype(self.dtype)
        q1 = 1 - 1.5 * q**2 * (1 - q / 2)
        q2 = 0.25 * (2 - q) ** 3
        return self._c(self._sigma) * (q1 * c1 + q2 * c2)

    def grad_w(self, r):
        q = np.asarray(r, dtype=self.dtype) * self._c(self._one_over_h)
        c1 = (1 - q >= 0).astype(self.dtype)
        c2 = ((2 - q < 1) & (2 - q >= 0)).astype(self.dtype)
        dq1 = -3.0 * q + 2.25 * q**2
        dq2 = -0.75 * (2 - q) ** 2
        return self._c(self._sigma) * self._c(self._one_over_h) * (dq1 * c1 + dq2 * c2)


class _WendlandHigh(_Base):
    """w = sigma * max(0, 1 - q/2)^n * poly(q); grad_w is what jax.grad gives (kernel.py:20-23):
    sigma / h * (-(n/2) q1^(n-1) poly + q1^n poly')."""

    _normalized_cutoff = 2.0

    def w(self, r):
        q = np.asarray(r, dtype=self.dtype) * self._c(self._one_over_h)
        q1 = np.maximum(self._c(0.0), self._c(1.0) - self._c(0.5) * q)
        return self._c(self._sigma) * (_ipow(q1, self._n) * self._poly(q))

    def grad_w(self, r):
        q = np.asarray(r, dtype=self.dtype) * self._c(self._one_over_h)
        q1 = np.maximum(self._c(0.0), self._c(1.0) - self._c(0.5) * q)
        inside = (q1 > 0).astype(self.dtype)  # d max(0, x)/dx
        dq = (self._c(-0.5 * self._n) * _ipow(q1, self._n - 1) * inside * self._poly(q)
              + _ipow(q1, self._n) * self._dpoly(q))
        return self._c(self._sigma) * self._c(self._one_over_h) * dq


class WendlandC4Kernel(_WendlandHigh):
    def __init__(self, h, dim=3, dtype=np.float64):
        super().__init__(h, dim, dtype)
        ooh = self._one_over_h
        self._sigma = {
            1: 3.0 / 4.0 * ooh,
            2: 9.0 / 4.0 / np.pi * ooh**2,
            3: 495.0 / 256.0 / np.pi * ooh**3,
        }[dim]
        self._n = 5 if dim == 1 else 6

    def _poly(self, q):
        if self.dim == 1:
            return 2.0 * q**2 + 2.5 * q + 1.0
        return self._c(35.0 / 12.0) * q**2 + 3 * q + 1.0

    def _dpoly(self, q):
        if self.dim == 1:
            return 4.0 * q + 2.5
        return self._c(35.0 / 6.0) * q + 3


class WendlandC6Kernel(_WendlandHigh):
    def __init__(self, h, dim=3, dtype=np.float64):
        super().__init__(h, dim, dtype)
        ooh = self._one_over_h
        self._sigma = {
            1: 55.0 / 64.0 * ooh,
            2: 78.0 / 28.0 / np.pi * ooh**2,
            3: 1365.0 / 512.0 / np.pi * ooh**3,
        }[dim]
        self._n = 7 if dim == 1 else 8

    def _poly(self, q):
        if self.dim == 1:
            return 21.0 / 8.0 * q**3 + 19.0 / 4.0 * q**2 + 3.5 * q + 1.0
        return 4.0 * q**3 + 6.25 * q**2 + 4 * q + 1.0

    def _dpoly(self, q):
        if self.dim == 1:
            return 63.0 / 8.0 * q**2 + 9.5 * q + 3.5
        return 12.0 * q**2 + 12.5 * q + 4


class GaussianKernel(_Base):
    _normalized_cutoff = 3.0

    def __init__(self, h, dim=3, dtype=np.float64):
        super().__init__(h, dim, dtype)
        self._sigma = 1.0 / np.pi ** (dim / 2) * self._one_over_h**dim

    def w(self, r):
        q = np.asarray(r, dtype=self.dtype) * self._c(self._one_over_h)
        return self._c(self._sigma) * (3 - q >= 0) * np.exp(-(q**2))

    def grad_w(self, r):
        q = np.asarray(r, dtype=self.dtype) * self._c(self._one_over_h)
        return (
            self._c(self._sigma)
            * self._c(self._one_over_h)
            * (3 - q >= 0)
            * (-2.0 * q)
            * np.exp(-(q**2))
        )


class SuperGaussianKernel(_Base):
    """kernel.py:185-201: sigma [q <= 3] exp(-q^2) (dim/2 + 1 - q^2)."""

    _normalized_cutoff = 3.0

    def __init__(self, h, dim=3, dtype=np.float64):
        super().__init__(h, dim, dtype)
        self._sigma = 1.0 / np.pi ** (dim / 2) * self._one_over_h**dim

    def w(self, r):
        q = np.asarray(r, dtype=self.dtype) * self._c(self._one_over_h)
        return (self._c(self._sigma) * (3 - q >= 0) * np.exp(-(q**2))
                * (self._c(self.dim / 2 + 1) - q**2))

    def grad_w(self, r):
        q = np.asarray(r, dtype=self.dtype) * self._c(self._one_over_h)
        # d/dq [e^{-q^2} (a - q^2)] = -2 q e^{-q^2} (a + 1 - q^2)
        return (self._c(self._sigma) * self._c(self._one_over_h) * (3 - q >= 0) * (-2.0 * q)
                * np.exp(-(q**2)) * (self._c(self.dim / 2 + 2) - q**2))


KERNELS = {
    "CSK": CubicKernel,
    "QSK": QuinticKernel,
    "WC2K": WendlandC2Kernel,
    "WC4K": WendlandC4Kernel,
    "WC6K": WendlandC6Kernel,
    "GK": GaussianKernel,
    "SGK": SuperGaussianKernel,
}
