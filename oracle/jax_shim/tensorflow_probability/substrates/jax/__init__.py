"""numpy stand-in for the few tensorflow_probability.substrates.jax symbols gpfy uses.

TEST INFRASTRUCTURE ONLY (see oracle/jax_shim/README.md).  Formulas follow TFP's published
definitions: Exp / Identity / FillTriangular bijectors, Normal / Bernoulli(probs=) log_prob,
math.interp_regular_1d_grid (linear, fill 'constant_extension').
"""
import math as _math
import types as _types

import numpy as _np


class _Bijector:
    pass


class _Exp(_Bijector):
    def forward(self, x):
        return _np.exp(x)

    def inverse(self, y):
        return _np.log(y)


class _Identity(_Bijector):
    def forward(self, x):
        return x

    def inverse(self, y):
        return y


def _fill_triangular(x):
    """TFP fill_triangular(upper=False): vector of n(n+1)/2 -> lower [n, n].

    concat([x[n:], reverse(x)]) reshaped to [n, n], then the lower band is kept, e.g.
    [1,2,3,4,5,6] -> [[4,0,0],[6,5,0],[3,2,1]] (the example in TFP's docstring).
    """
    x = _np.asarray(x)
    m = x.shape[-1]
    n = int(round((_math.sqrt(1 + 8 * m) - 1) / 2))
    if n * (n + 1) // 2 != m:
        raise ValueError("fill_triangular: bad length")
    xc = _np.concatenate([x[..., n:], x[..., ::-1]], axis=-1)
    return _np.tril(xc.reshape(x.shape[:-1] + (n, n)))


def _fill_triangular_inverse(y):
    """Index-based inverse of _fill_triangular (reads the lower triangle only)."""
    y = _np.asarray(y)
    n = y.shape[-1]
    m = n * (n + 1) // 2
    idx = _fill_triangular(_np.arange(1, m + 1, dtype=_np.int64))
    out = _np.zeros(y.shape[:-2] + (m,), dtype=y.dtype)
    ii, jj = _np.nonzero(idx)
    out[..., idx[ii, jj] - 1] = y[..., ii, jj]
    return out


class _FillTriangular(_Bijector):
    def forward(self, x):
        return _fill_triangular(x)

    def inverse(self, y):
        return _fill_triangular_inverse(y)


bijectors = _types.SimpleNamespace(Bijector=_Bijector, Exp=_Exp, Identity=_Identity, FillTriangular=_FillTriangular)


class _Normal:
    def __init__(self, loc, scale):
        self.loc, self.scale = loc, scale

    def log_prob(self, x):
        z = (x - self.loc) / self.scale
        return -0.5 * z * z - _np.log(self.scale) - 0.5 * _math.log(2.0 * _math.pi)


class _Bernoulli:
    def __init__(self, logits=None, probs=None):
        assert probs is not None
        self.probs = probs

    def log_prob(self, event):
        event = _np.asarray(event, dtype=_np.float64)
        lp0 = _np.log1p(-self.probs)
        lp1 = _np.log(self.probs)
        return _np.where(event == 1, 0.0, lp0 * (1 - event)) + _np.where(event == 0, 0.0, lp1 * event)


distributions = _types.SimpleNamespace(Normal=_Normal, Bernoulli=_Bernoulli, Distribution=object)


def _interp_regular_1d_grid(x, x_ref_min, x_ref_max, y_ref, axis=-1, **_k):
    x = _np.asarray(x, dtype=_np.float64)
    y_ref = _np.asarray(y_ref)
    nd = y_ref.shape[-1]
    xi = (x - x_ref_min) / (x_ref_max - x_ref_min) * (nd - 1)
    xi = _np.clip(xi, 0.0, nd - 1.0)
    lo = _np.floor(xi)
    hi = _np.minimum(lo + 1, nd - 1)
    t = xi - lo
    return (1 - t) * y_ref[..., lo.astype(int)] + t * y_ref[..., hi.astype(int)]


math = _types.SimpleNamespace(interp_regular_1d_grid=_interp_regular_1d_grid)
