"""Gegenbauer polynomials (mirror of gpfy/gegenbauer.py).

`gegenbauer(level, alpha, x)` evaluates on the GPU for array inputs (C-ABI `gpfy_gegenbauer`);
`gegenbauer_host` is the same recurrence for the N-independent host-side parameter algebra
(Gram matrices of the fundamental points, C_n(1), the Funk-Hecke nodes).
"""
import numpy as np

from . import ops


def gegenbauer_host(level: int, alpha: float, x):
    """gegenbauer.py:27-60 in numpy (same association order as :49)."""
    x = np.asarray(x, dtype=np.float64)
    level = int(level)
    if level == 0:
        return np.ones_like(x)
    C_prev, C = np.ones_like(x), 2 * alpha * x
    n = 2.0
    while n <= level:
        C, C_prev = (2 * x * (n + alpha - 1) * C - (n + 2 * alpha - 2) * C_prev) / n, C
        n += 1.0
    return C


def gegenbauer_grad_host(level: int, alpha: float, x):
    """gegenbauer.py:70-75: 2 alpha C_{n-1}^{alpha+1}(x); zeros at level 0."""
    x = np.asarray(x, dtype=np.float64)
    if int(level) == 0:
        return np.zeros_like(x)
    return 2 * alpha * gegenbauer_host(int(level) - 1, alpha + 1, x)


def gegenbauer(level, alpha, x):
    """C_level^alpha(x) on the device; returns a CUDA tensor shaped like x (gegenbauer.py:63-66)."""
    return ops.gegenbauer(int(level), float(alpha), x)


def gegenbauer_fwd(level, alpha, x):
    """(value, derivative) pair of the reference's custom VJP (gegenbauer.py:70-75)."""
    return ops.gegenbauer(int(level), float(alpha), x, with_grad=True)


class GegenbauerLookupTable:
    """gegenbauer.py:88-147.  The reference uses the table only for C_n(1) (exact at the last grid
    node) and for PolynomialDecay's shape function; `__call__` interpolates on the host (init-time algebra), the device
    kernels re-evaluate the two bracketing grid nodes instead of storing the table (csrc/kmat.cu)."""

    def __init__(self, max_level: int, alpha: float, grid_resolution: int = 100_000):
        self.max_level, self.alpha, self.grid_resolution = int(max_level), float(alpha), int(grid_resolution)
        self._grid = None

    def _ensure(self):
        if self._grid is None:
            x = np.linspace(-1, 1, self.grid_resolution)
            self._grid = np.stack([gegenbauer_host(n, self.alpha, x) for n in range(self.max_level + 1)])
        return self._grid

    def at_one(self, level: int):
        return gegenbauer_host(level, self.alpha, np.array(1.0))

    def __call__(self, level, alpha, x):
        g = self._ensure()[int(level)]
        x = np.asarray(x, dtype=np.float64)
        nd = g.shape[-1]
        xi = np.clip((x + 1.0) / 2.0 * (nd - 1), 0.0, nd - 1.0)
        lo = np.floor(xi)
        hi = np.minimum(lo + 1, nd - 1)
        t = xi - lo
        return (1 - t) * g[lo.astype(int)] + t * g[hi.astype(int)]
