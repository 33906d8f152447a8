"""numpy-backed stand-in for the subset of the JAX API used by /root/reference/src/gpfy.

TEST INFRASTRUCTURE ONLY (see ../README.md).  Eager float64 numpy; forward evaluation only.
"""
import functools

import numpy as _np

from . import tree_util  # noqa: F401
from . import tree  # noqa: F401
from . import lax  # noqa: F401
from . import random  # noqa: F401
from . import numpy  # noqa: F401
from . import scipy  # noqa: F401
from . import debug  # noqa: F401

Array = _np.ndarray


class _At:
    """`x.at[idx].set(v)` of jax arrays (functional update), used by gpfy/spherical.py:341."""

    def __init__(self, arr):
        self._arr = arr
        self._idx = None

    def __getitem__(self, idx):
        self._idx = idx
        return self

    def set(self, value):
        out = _np.array(self._arr, copy=True)
        out[self._idx] = value
        return out.view(_ShimArray)


class _ShimArray(_np.ndarray):
    """ndarray with the `.at` property; only what `vmap` returns needs it."""

    @property
    def at(self):
        return _At(self)


class _Config:
    def update(self, *_a, **_k):
        return None


config = _Config()


def jit(fun=None, **_kw):
    if fun is None:
        return lambda f: f
    return fun


def vmap(fun, in_axes=0, out_axes=0):
    """Eager vmap over the leading axis of every positional argument."""

    def mapped(*args):
        n = None
        for a in tree_util.tree_leaves(args):
            n = _np.shape(a)[0]
            break
        outs = []
        for i in range(n):
            sl = tree_util.tree_map(lambda a: _np.asarray(a)[i], args)
            outs.append(fun(*sl))
        return tree_util.tree_map(lambda *xs: _np.stack([_np.asarray(x) for x in xs], 0).view(_ShimArray), *outs)

    return mapped


class custom_vjp:
    """Forward-only: the decorated function is called directly; defvjp is recorded, never used."""

    def __init__(self, fun, nondiff_argnums=()):
        self.fun = fun
        functools.update_wrapper(self, fun)

    def defvjp(self, fwd, bwd):
        self.fwd, self.bwd = fwd, bwd

    def __call__(self, *args, **kwargs):
        return self.fun(*args, **kwargs)


def grad(*_a, **_k):
    raise NotImplementedError("jax_shim is forward-only; gradients come from oracle/torch_oracle.py")


value_and_grad = grad
