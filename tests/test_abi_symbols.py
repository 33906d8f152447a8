"""The C-ABI library loads on a CPU-only box and exports every symbol include/gpfy_b200.h declares.
Host-only calls (descriptor validation, layouts, workspace sizes) are exercised; nothing computes."""
import ctypes
import os
import re

import pytest

from gpfy_b200 import _lib
from gpfy_b200.build import LIB, build_library

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "gpfy_b200.h")).read()
    return sorted(set(re.findall(r"GPFY_API\s+[\w\s\*]+?\b(gpfy_\w+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    build_library()
    assert os.path.exists(LIB)
    lib = ctypes.CDLL(LIB)
    names = _declared()
    assert len(names) >= 17
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/gpfy_b200.h but not exported"
    # and the ctypes table binds exactly the declared entry points (+ debug helpers)
    bound = set(_lib.exported_symbols())
    assert set(names) <= bound
    assert all(b in names or b.startswith("gpfy_debug_") for b in bound)


def test_descriptor_validation_and_layout():
    lib = _lib.load()
    m = [1, 9] + [30] * 9
    d = _lib.make_desc(8, m)
    assert lib.gpfy_num_inducing(ctypes.byref(d)) == 280
    lay = _lib.GpfyElboLayout()
    _lib.check(lib.gpfy_elbo_packed_layout(ctypes.byref(d), ctypes.byref(lay)))
    M, F, dim = 280, 11, 9
    assert lay.data_term == 0 and lay.d_mu == 1 and lay.d_sigma == 1 + M
    assert lay.total == 1 + M + M * M + F + M * dim + (1 + 81 + 9 * 900) + 8 + 3
    nbytes = ctypes.c_size_t(0)
    for op in (_lib.OP_FEATURES, _lib.OP_ELBO, _lib.OP_PREDICT):
        _lib.check(lib.gpfy_workspace_bytes(ctypes.byref(d), 1000, op, ctypes.byref(nbytes)))
        assert nbytes.value > 0 and nbytes.value % 256 == 0
    # error behaviour: negative code + message, no exception crosses the boundary
    bad = _lib.make_desc(8, m)
    bad.abi_version = 99
    assert lib.gpfy_num_inducing(ctypes.byref(bad)) == -1
    assert b"version" in lib.gpfy_last_error_string()
    bad = _lib.make_desc(40, [1, 2])
    assert lib.gpfy_num_inducing(ctypes.byref(bad)) == -2  # GPFY_EUNSUPPORTED
    with pytest.raises(_lib.GpfyError):
        _lib.check(lib.gpfy_workspace_bytes(ctypes.byref(d), -5, 1, ctypes.byref(nbytes)))


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    from gpfy_b200.ops import HotPath
    with pytest.raises(_lib.GpfyError):
        HotPath(8, [1, 9, 30])
