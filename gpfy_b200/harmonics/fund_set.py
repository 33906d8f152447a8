"""Loader of the precomputed fundamental systems (mirror of gpfy/harmonics/fund_set.py:22-39).

The `fs_{dim}D.npz` files are DATA shipped with the gpfy package (not source); they are looked up,
in order, in $GPFY_FUNDAMENTAL_SYSTEM_DIR, in an installed `gpfy` package, and in the reference
checkout.  Same error behaviour as the reference: ValueError for a missing dimension or degree.
"""
import importlib.util
import os
from pathlib import Path
from typing import Callable

import numpy as np


def _search_dirs():
    env = os.environ.get("GPFY_FUNDAMENTAL_SYSTEM_DIR")
    if env:
        yield Path(env)
    spec = importlib.util.find_spec("gpfy")
    if spec is not None and spec.submodule_search_locations:
        for loc in spec.submodule_search_locations:
            yield Path(loc) / "fundamental_system"
    yield Path("/root/reference/src/gpfy/fundamental_system")


def fundamental_set_loader(dimension: int) -> Callable[[int], np.ndarray]:
    cache = None
    for d in _search_dirs():
        f = d / f"fs_{dimension}D.npz"
        if f.exists():
            with np.load(f) as z:
                cache = {k: v for k, v in z.items()}
            break
    if cache is None:
        raise ValueError(f"Fundamental system for dimension {dimension} has not been precomputed.")

    def load(degree: int) -> np.ndarray:
        key = f"degree_{degree}"
        if key not in cache:
            raise ValueError()
        return cache[key]

    return load
