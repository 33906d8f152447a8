"""numpy PCG64 stand-in for jax.random (NOT bit-compatible with threefry)."""
import numpy as _np

PRNGKeyArray = _np.ndarray


def PRNGKey(seed):
    return _np.array([0, int(seed)], dtype=_np.uint32)


key = PRNGKey


def split(key, num=2):
    ss = _np.random.SeedSequence([int(k) for k in _np.asarray(key).ravel()])
    return [_np.array(c.generate_state(2), dtype=_np.uint32) for c in ss.spawn(num)]


def normal(key, shape=(), dtype=_np.float64):
    rng = _np.random.default_rng([int(k) for k in _np.asarray(key).ravel()])
    return rng.standard_normal(shape).astype(dtype)
