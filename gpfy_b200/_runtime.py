"""Process-wide runtime helpers: hot-path cache, RNG keys, errors."""
import numpy as np

from ._lib import GpfyError  # noqa: F401

_CACHE = {}


def hot_path(input_dim, m, num_outputs=1, kernel_order=1, likelihood="gaussian", q_kind="tril", weight_mode="ard", projection_dim=0):
    """One `ops.HotPath` (descriptor + scratch) per static configuration."""
    from .ops import HotPath, require_cuda
    dev = require_cuda()
    key = (int(input_dim), tuple(int(v) for v in m), int(num_outputs), int(kernel_order), likelihood, q_kind, weight_mode, int(projection_dim), str(dev))
    hp = _CACHE.get(key)
    if hp is None:
        hp = HotPath(input_dim, m, num_outputs, kernel_order, likelihood, q_kind, weight_mode, device=dev, projection_dim=projection_dim)
        _CACHE[key] = hp
    return hp


def as_rng(key):
    """The reference takes a jax PRNG key; here an int seed, a numpy Generator or a (2,) uint32 key."""
    if isinstance(key, np.random.Generator):
        return key
    return np.random.default_rng([int(k) for k in np.atleast_1d(np.asarray(key)).ravel()])


def cuda_available():
    import torch
    return torch.cuda.is_available()
