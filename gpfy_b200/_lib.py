"""ctypes binding of the C-ABI in include/gpfy_b200.h.  No fallback: a missing library is an error."""
import ctypes
import os

from .build import LIB

GPFY_ABI_VERSION = 2
GPFY_MAX_DEGREES = 32
LIK_GAUSSIAN, LIK_BERNOULLI = 0, 1
Q_TRIL, Q_FULLCOV = 0, 1
W_ARD, W_SCALAR, W_PROJECTION = 0, 1, 2
OP_FEATURES, OP_ELBO, OP_PREDICT, OP_FEATURES_VJP = 0, 1, 2, 3
PREC_F64, PREC_TF32X3, PREC_F64_I8 = 0, 1, 2


class GpfyDesc(ctypes.Structure):
    _fields_ = [
        ("struct_size", ctypes.c_uint32),
        ("abi_version", ctypes.c_uint32),
        ("input_dim", ctypes.c_int32),
        ("num_degrees", ctypes.c_int32),
        ("m", ctypes.c_int32 * GPFY_MAX_DEGREES),
        ("num_outputs", ctypes.c_int32),
        ("kernel_order", ctypes.c_int32),
        ("likelihood", ctypes.c_int32),
        ("q_kind", ctypes.c_int32),
        ("weight_mode", ctypes.c_int32),
        ("projection_dim", ctypes.c_int32),
        ("precision", ctypes.c_int32),
        ("alpha", ctypes.c_double),
    ]


class GpfyElboLayout(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int64) for n in
                ("total", "data_term", "d_mu", "d_sigma", "d_eig", "d_vhat", "d_lcat", "d_w", "d_b", "d_v", "d_sigma2")]


class GpfyStepLayout(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in ("kernel_kind", "truncation_level", "w_identity", "n_w")] + \
               [(n, ctypes.c_int64) for n in ("th_w", "th_b", "th_v", "th_beta", "th_s2", "th_mu", "th_sig", "th_V", "th_total",
                                              "c_w", "c_b", "c_v", "c_beta", "c_eig", "c_deig", "c_s2", "c_mu", "c_sig", "c_V", "c_total")]


MAX_SHAPE_LEVELS = 64
SHAPE_ARCCOSINE, SHAPE_NTK, SHAPE_POLYDECAY = 0, 1, 2


class GpfyShapeFn(ctypes.Structure):
    _fields_ = [("struct_size", ctypes.c_uint32), ("kind", ctypes.c_int32), ("depth", ctypes.c_int32), ("order", ctypes.c_int32),
                ("num_levels", ctypes.c_int32), ("grid_resolution", ctypes.c_int32), ("coef", ctypes.c_double * MAX_SHAPE_LEVELS)]


def make_shape_fn(kind, depth=1, order=1, coef=(), grid_resolution=100_000):
    if len(coef) > MAX_SHAPE_LEVELS:
        raise GpfyError(f"truncation level {len(coef)} exceeds {MAX_SHAPE_LEVELS}")
    fn = GpfyShapeFn()
    fn.struct_size = ctypes.sizeof(GpfyShapeFn)
    fn.kind, fn.depth, fn.order = int(kind), int(depth), int(order)
    fn.num_levels, fn.grid_resolution = len(coef), int(grid_resolution)
    for i, c in enumerate(coef):
        fn.coef[i] = float(c)
    return fn


class GpfyError(RuntimeError):
    pass


_P = ctypes.c_void_p
_SIGNATURES = {
    "gpfy_abi_version": (ctypes.c_int, []),
    "gpfy_last_error_string": (ctypes.c_char_p, []),
    "gpfy_debug_launch_count": (ctypes.c_ulonglong, []),
    "gpfy_num_inducing": (ctypes.c_int64, [ctypes.POINTER(GpfyDesc)]),
    "gpfy_elbo_packed_layout": (ctypes.c_int, [ctypes.POINTER(GpfyDesc), ctypes.POINTER(GpfyElboLayout)]),
    "gpfy_workspace_bytes": (ctypes.c_int, [ctypes.POINTER(GpfyDesc), ctypes.c_int64, ctypes.c_int,
                                            ctypes.POINTER(ctypes.c_size_t)]),
    "gpfy_gegenbauer": (ctypes.c_int, [_P, ctypes.c_int, ctypes.c_double, _P, ctypes.c_int64, _P, _P]),
    "gpfy_sh_features_fwd": (ctypes.c_int, [_P, ctypes.POINTER(GpfyDesc), ctypes.c_int64, _P, ctypes.c_int, _P, _P, _P, _P,
                                            _P, ctypes.c_int, _P, _P, _P, ctypes.c_size_t]),
    "gpfy_sh_features_vjp": (ctypes.c_int, [_P, ctypes.POINTER(GpfyDesc), ctypes.c_int64, _P, ctypes.c_int, _P, _P, _P, _P,
                                            _P, ctypes.c_int, _P, _P, _P, _P, _P, _P, _P, ctypes.c_size_t]),
    "gpfy_svgp_elbo_valgrad": (ctypes.c_int, [_P, ctypes.POINTER(GpfyDesc), ctypes.c_int64] + [_P] * 12 +
                               [_P, ctypes.c_size_t]),
    "gpfy_svgp_elbo_begin": (ctypes.c_int, [_P, ctypes.POINTER(GpfyDesc), ctypes.c_int64] + [_P] * 8 + [_P, ctypes.c_size_t]),
    "gpfy_svgp_elbo_rows": (ctypes.c_int, [_P, ctypes.POINTER(GpfyDesc), ctypes.c_int64, ctypes.c_int64] + [_P] * 4 +
                            [_P, ctypes.c_size_t]),
    "gpfy_svgp_elbo_finish": (ctypes.c_int, [_P, ctypes.POINTER(GpfyDesc), ctypes.c_int64] + [_P] * 6 + [_P, ctypes.c_size_t]),
    "gpfy_inducing_basis_fwd": (ctypes.c_int, [_P, ctypes.POINTER(GpfyDesc), _P, ctypes.c_int, _P, _P]),
    "gpfy_inducing_basis_vjp": (ctypes.c_int, [_P, ctypes.POINTER(GpfyDesc), _P, _P, ctypes.c_int, _P, _P, _P]),
    "gpfy_inducing_basis_fwd_ws": (ctypes.c_int, [_P, ctypes.POINTER(GpfyDesc), _P, ctypes.c_int, _P, _P, _P, ctypes.c_size_t]),
    "gpfy_inducing_basis_vjp_ws": (ctypes.c_int, [_P, ctypes.POINTER(GpfyDesc), _P, _P, ctypes.c_int, _P, _P, _P, _P, ctypes.c_size_t]),
    "gpfy_step_constrain": (ctypes.c_int, [_P, ctypes.POINTER(GpfyDesc), ctypes.POINTER(GpfyStepLayout), _P, _P, _P, _P]),
    "gpfy_step_grad": (ctypes.c_int, [_P, ctypes.POINTER(GpfyDesc), ctypes.POINTER(GpfyStepLayout), _P, _P, _P, _P, _P, ctypes.c_double, _P, _P]),
    "gpfy_step_adam": (ctypes.c_int, [_P, ctypes.c_int64, _P, _P, _P, _P, _P] + [ctypes.c_double] * 4 + [ctypes.c_int64]),
    "gpfy_prior_kl_valgrad": (ctypes.c_int, [_P, ctypes.POINTER(GpfyDesc), _P, _P, _P]),
    "gpfy_prior_kl_valgrad_ws": (ctypes.c_int, [_P, ctypes.POINTER(GpfyDesc), _P, _P, _P, _P, ctypes.c_size_t]),
    "gpfy_predict_diag": (ctypes.c_int, [_P, ctypes.POINTER(GpfyDesc), ctypes.c_int64] + [_P] * 11 + [_P, ctypes.c_size_t]),
    "gpfy_allreduce_packed": (ctypes.c_int, [_P, _P, _P, ctypes.c_int64]),
    "gpfy_debug_kernel_timing": (ctypes.c_int, [ctypes.c_int]),
    "gpfy_debug_i8_plan": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.c_int]),
    "gpfy_debug_features3_plan": (ctypes.c_int, [ctypes.POINTER(ctypes.c_int), ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.c_int]),
    "gpfy_debug_set_option": (ctypes.c_int, [ctypes.c_int, ctypes.c_int]),
    "gpfy_debug_kernel_time": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int64)]),
    "gpfy_to_sphere": (ctypes.c_int, [_P, ctypes.POINTER(GpfyDesc), ctypes.c_int64, _P, _P, _P, _P, _P]),
    "gpfy_variational_expectations": (ctypes.c_int, [_P, ctypes.POINTER(GpfyDesc), ctypes.c_int64, _P, _P, _P, _P, _P]),
    "gpfy_project_q": (ctypes.c_int, [_P, ctypes.POINTER(GpfyDesc), ctypes.c_int64, _P, _P, _P, _P, _P, _P, ctypes.c_size_t]),
    "gpfy_minibatch_gather": (ctypes.c_int, [_P, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, _P,
                                             ctypes.c_int, _P, ctypes.c_int, _P, _P, _P]),
    "gpfy_shuffle_index": (ctypes.c_int, [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int64, ctypes.c_int64, ctypes.POINTER(ctypes.c_int64)]),
    "gpfy_step_adam_state": (ctypes.c_int, [_P, ctypes.c_int64, _P, _P, _P, _P, _P, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                            ctypes.c_double, _P]),
    "gpfy_minibatch_gather_state": (ctypes.c_int, [_P, ctypes.c_uint64, ctypes.c_int64, ctypes.c_int64, _P, _P, ctypes.c_int, _P, ctypes.c_int,
                                                   _P, _P, _P]),
    "gpfy_step_advance": (ctypes.c_int, [_P, _P, ctypes.c_int64, ctypes.c_int64, _P, _P, ctypes.c_int64]),
    "gpfy_chunk_rows": (ctypes.c_int, [ctypes.POINTER(GpfyDesc), ctypes.c_int64, ctypes.POINTER(ctypes.c_int64)]),
    "gpfy_contraction_path": (ctypes.c_int, [ctypes.POINTER(GpfyDesc), ctypes.c_int64, ctypes.c_int, ctypes.POINTER(ctypes.c_int32)]),
    "gpfy_kmat_workspace_bytes": (ctypes.c_int, [ctypes.POINTER(GpfyDesc), ctypes.c_int64, ctypes.c_int64, ctypes.POINTER(ctypes.c_size_t)]),
    "gpfy_shape_function": (ctypes.c_int, [_P, ctypes.POINTER(GpfyShapeFn), ctypes.c_double, _P, ctypes.c_int64, _P]),
    "gpfy_kernel_matrix": (ctypes.c_int, [_P, ctypes.POINTER(GpfyDesc), ctypes.POINTER(GpfyShapeFn), ctypes.c_int64, _P, ctypes.c_int64, _P,
                                          _P, _P, _P, ctypes.c_int, _P, _P, ctypes.c_size_t]),
    "gpfy_project_variance": (ctypes.c_int, [_P, ctypes.POINTER(GpfyDesc), ctypes.c_int64, _P, _P, ctypes.c_int, _P, _P, _P, ctypes.c_size_t]),
}

_lib = None


def exported_symbols():
    """Names every entry point include/gpfy_b200.h declares (plus the launch counter)."""
    return sorted(_SIGNATURES)


def load():
    """Loads gpfy_b200/csrc/libgpfy_b200.so; raises if it has not been built (no CPU fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB):
        raise GpfyError(f"{LIB} is missing: run `python -m gpfy_b200.build` (nvcc, sm_100a). "
                        "gpfy_b200 has no CPU or PyTorch fallback.")
    lib = ctypes.CDLL(LIB, mode=ctypes.RTLD_GLOBAL)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.gpfy_abi_version() != GPFY_ABI_VERSION:
        raise GpfyError("ABI version mismatch between gpfy_b200/_lib.py and libgpfy_b200.so")
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise GpfyError(f"gpfy_b200 error {rc}: {load().gpfy_last_error_string().decode()}")


def make_desc(input_dim, m, num_outputs=1, kernel_order=1, likelihood=LIK_GAUSSIAN, q_kind=Q_TRIL, weight_mode=W_ARD,
              projection_dim=0, precision=PREC_F64):
    d = GpfyDesc()
    d.struct_size = ctypes.sizeof(GpfyDesc)
    d.abi_version = GPFY_ABI_VERSION
    d.input_dim = int(input_dim)
    d.num_degrees = len(m)
    if len(m) > GPFY_MAX_DEGREES:
        raise GpfyError(f"num_frequencies {len(m)} > {GPFY_MAX_DEGREES}")
    for i, mi in enumerate(m):
        d.m[i] = int(mi)
    d.num_outputs = int(num_outputs)
    d.kernel_order = int(kernel_order)
    d.likelihood = int(likelihood)
    d.q_kind = int(q_kind)
    d.weight_mode = int(weight_mode)
    d.projection_dim = int(projection_dim) if weight_mode == W_PROJECTION else 0
    d.precision = int(precision)
    sphere_dim = d.projection_dim if d.projection_dim > 0 else int(input_dim)
    d.alpha = (sphere_dim + 1 - 2.0) / 2.0
    return d
