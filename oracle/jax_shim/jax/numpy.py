"""jax.numpy stand-in: numpy float64 with the few signature differences patched."""
import numpy as _np
from numpy import *  # noqa: F401,F403
from numpy import linalg  # noqa: F401

float64 = _np.float64
int32 = _np.int32
pi = _np.pi
nan = _np.nan


def array(x, dtype=None, **_k):
    return _np.array(x, dtype=dtype)


asarray = array


def _ax(axis):
    return tuple(axis) if isinstance(axis, list) else axis


def sum(a, axis=None, **kw):  # noqa: A001
    return _np.sum(a, axis=_ax(axis), **kw)


def mean(a, axis=None, **kw):
    return _np.mean(a, axis=_ax(axis), **kw)


def ones_like(a, dtype=None, **_k):
    return _np.ones_like(a, dtype=dtype)


def zeros_like(a, dtype=None, **_k):
    return _np.zeros_like(a, dtype=dtype)
