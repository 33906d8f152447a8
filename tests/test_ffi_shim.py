"""The XLA typed-FFI shim (gpfy_b200/csrc/xla_ffi_shim.cc) cannot be built against jaxlib here; it is at least COMPILED:
type-checked against tests/ffi_stub (a restatement of the public FFI C++ API) so that every handler's signature matches its
binding chain and every C-ABI call it forwards to matches include/gpfy_b200.h.  Also: one handler per C-ABI entry a JAX front
end needs, and the attribute block carries projection_dim and precision."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "gpfy_b200", "csrc", "xla_ffi_shim.cc")


def test_shim_type_checks_against_the_ffi_api_stub():
    cmd = ["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Werror", "-x", "c++", "-I", os.path.join(ROOT, "tests", "ffi_stub"),
           "-I", "/usr/local/cuda/include", SHIM]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-4000:]


def test_shim_covers_the_entries_a_jax_front_end_needs():
    src = open(SHIM).read()
    handlers = set(re.findall(r"XLA_FFI_DEFINE_HANDLER_SYMBOL\((gpfy_ffi_\w+)", src))
    want = {"gpfy_ffi_sh_features_fwd", "gpfy_ffi_sh_features_vjp", "gpfy_ffi_svgp_elbo_valgrad", "gpfy_ffi_predict_diag",
            "gpfy_ffi_gegenbauer", "gpfy_ffi_to_sphere", "gpfy_ffi_variational_expectations", "gpfy_ffi_project_q",
            "gpfy_ffi_prior_kl_valgrad", "gpfy_ffi_inducing_basis_fwd", "gpfy_ffi_inducing_basis_vjp", "gpfy_ffi_project_variance"}
    assert want <= handlers, want - handlers
    for attr in ("projection_dim", "precision", "weight_mode", "q_kind", "likelihood", "kernel_order", "num_outputs", "input_dim"):
        assert f'Attr<int32_t>("{attr}")' in src, attr
    # every forwarded call is a declared C-ABI entry
    header = open(os.path.join(ROOT, "include", "gpfy_b200.h")).read()
    for call in set(re.findall(r"status\((gpfy_\w+)\(", src)):
        assert re.search(r"GPFY_API int " + call + r"\(", header), call


def test_integration_md_packed_layout_matches_the_library():
    """The `_layout` helper printed in INTEGRATION.md (what a GPfY maintainer pastes into gpfy/_b200.py to unpack the packed buffer
    in `_bwd`) is executed here and compared with gpfy_elbo_packed_layout for ARD, scalar and projection weights."""
    import ctypes
    from gpfy_b200 import _lib
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    start = text.index("def _layout(static):")
    end = text.index("f64 = lambda", start)
    ns = {}
    exec(text[start:end], ns)
    lib = _lib.load()
    for m, d, P, wm, pd in (([1, 9, 30, 30], 8, 1, 0, 0), ([1, 4, 9], 3, 2, 1, 0), ([1, 4, 6], 6, 1, 2, 3)):
        static = (tuple(m), d, P, 1, 0, 0, wm, pd, 0, 0.0)
        off, total = ns["_layout"](static)
        desc = _lib.make_desc(d, m, P, 1, _lib.LIK_GAUSSIAN, _lib.Q_TRIL, wm, pd)
        lay = _lib.GpfyElboLayout()
        _lib.check(lib.gpfy_elbo_packed_layout(ctypes.byref(desc), ctypes.byref(lay)))
        assert total == lay.total
        for k, name in (("data_term", "data_term"), ("mu", "d_mu"), ("sigma", "d_sigma"), ("eig", "d_eig"), ("vhat", "d_vhat"),
                        ("lcat", "d_lcat"), ("w", "d_w"), ("b", "d_b"), ("v", "d_v"), ("sigma2", "d_sigma2")):
            assert off[k][0] == getattr(lay, name), (k, off[k], getattr(lay, name))
