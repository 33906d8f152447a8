"""Loader of the precomputed fundamental systems (mirror of gpfy/harmonics/fund_set.py:22-39).

The tables are DATA (unit vectors per dimension and degree) that the gpfy package ships; here they live
repacked in gpfy_b200/data/fundamental_systems.npz (tools/pack_fundamental_systems.py), keys
d{dim}_n{degree}.  $GPFY_FUNDAMENTAL_SYSTEM_DIR, if set, points at a directory of the original
fs_{dim}D.npz files instead.  Same error behaviour as the reference: ValueError for a missing dimension
(:31) or degree (:36)."""
import os
from pathlib import Path
from typing import Callable

import numpy as np

_PACKED = Path(__file__).resolve().parent.parent / "data" / "fundamental_systems.npz"
_cache = {}


def _tables(dimension: int):
    if dimension in _cache:
        return _cache[dimension]
    tab = None
    env = os.environ.get("GPFY_FUNDAMENTAL_SYSTEM_DIR")
    if env and (Path(env) / f"fs_{dimension}D.npz").exists():
        with np.load(Path(env) / f"fs_{dimension}D.npz") as z:
            tab = {int(k.split("_")[1]): z[k] for k in z.files}
    elif _PACKED.exists():
        with np.load(_PACKED) as z:
            pre = f"d{dimension}_n"
            tab = {int(k[len(pre):]): z[k] for k in z.files if k.startswith(pre)} or None
    _cache[dimension] = tab
    return tab


def fundamental_set_loader(dimension: int) -> Callable[[int], np.ndarray]:
    tab = _tables(int(dimension))
    if tab is None:
        raise ValueError(f"Fundamental system for dimension {dimension} has not been precomputed.")

    def load(degree: int) -> np.ndarray:
        if int(degree) not in tab:
            raise ValueError(f"Degree {degree} is not part of the precomputed fundamental system of dimension {dimension}.")
        return tab[int(degree)]

    return load
