"""Host-side planning of the barrier-free feature kernel (csrc/features3.cuh ft3_groups / ft3_smem_bytes): the degree groups are
contiguous, cover every degree once and are balanced around ~96 rows; shapes with a degree of more than 32 rows are refused (they
stay on features2_kernel); the staged lower triangles of Linv are counted as the kernel lays them out.  No GPU involved."""
import ctypes
from math import comb

import pytest

from gpfy_b200 import _lib

MAX_DEGREES = 32


def _plan(m, Mp, dimp):
    arr = (ctypes.c_int * len(m))(*m)
    out = (ctypes.c_int * (6 + MAX_DEGREES + 1))()
    assert _lib.load().gpfy_debug_features3_plan(arr, len(m), Mp, dimp, out, len(out)) == 0
    if not out[0]:
        return None
    n = out[1]
    return {"ngrp": n, "linv": out[2], "smem_l": out[3], "smem": out[4], "lo": [out[5 + g] for g in range(n + 1)]}


def _degrees(dim, F, T):
    """m_l = min(N(dim, l), T): N(dim, l) harmonics of degree l on S^(dim-1) (spherical_harmonics.py:31-47)."""
    n = lambda l: 1 if l == 0 else (dim if l == 1 else comb(l + dim - 1, l) - comb(l + dim - 3, l - 2))
    return [min(n(l), T) for l in range(F)]


@pytest.mark.parametrize("dim,F,T", [(9, 11, 30), (9, 13, 30), (3, 9, 30), (17, 6, 30), (9, 11, 20), (32, 4, 32), (3, 32, 32), (9, 1, 30), (9, 2, 30), (3, 12, 12)])
def test_groups_cover_the_degrees_and_are_balanced(dim, F, T):
    m = _degrees(dim, F, T)
    Mp = (sum(m) + 31) // 32 * 32
    dimp = (dim + 3) // 4 * 4
    p = _plan(m, Mp, dimp)
    assert p is not None
    lo = p["lo"]
    assert lo[0] == 0 and lo[-1] == F and all(a < b for a, b in zip(lo, lo[1:]))          # contiguous, non-empty, complete
    rows = [sum(m[a:b]) for a, b in zip(lo, lo[1:])]
    total = sum(m)
    assert p["ngrp"] == max(1, (total + 48) // 96)
    if p["ngrp"] > 1:
        target = total / p["ngrp"]
        assert max(rows) <= target + 32 and min(rows) >= target - 32, rows               # within one degree of the mean
    # the staged triangles: row tile i of a degree is [8][8 i + 12]
    tri = lambda nt: sum(8 * (8 * i + 12) for i in range(nt))
    assert p["linv"] == sum(tri((ml + 7) // 8) for ml in m)
    assert p["smem_l"] - p["smem"] == 8 * p["linv"]
    assert p["smem"] >= 8 * (Mp * dimp + 16 * (32 * 32 + 32))


def test_headline_shape_stages_the_triangles_and_13_degrees_do_not():
    p = _plan(_degrees(9, 11, 30), 288, 12)
    assert [b - a for a, b in zip(p["lo"], p["lo"][1:])] == [5, 3, 3] and p["smem_l"] <= 227 * 1024 - 64
    p = _plan(_degrees(9, 13, 30), 352, 12)
    assert p["ngrp"] == 4 and p["smem"] <= 227 * 1024 - 64 < p["smem_l"]


@pytest.mark.parametrize("m", [[1, 9, 45, 64], [1, 33], [64]])
def test_degrees_beyond_32_rows_are_refused(m):
    assert _plan(m, (sum(m) + 31) // 32 * 32, 12) is None
