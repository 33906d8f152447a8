"""Device-side operators of the hot path: thin Python over the C-ABI.

torch is used only for device memory and the current CUDA stream; every computation happens in
libgpfy_b200.so.  All arrays are float64 (the reference forces x64, utils.py:26 / param.py:151).
"""
import ctypes

import numpy as np
import torch

from . import _lib


def _as_dev(x, device):
    if torch.is_tensor(x):
        t = x.to(device=device, dtype=torch.float64)
    else:
        t = torch.as_tensor(np.ascontiguousarray(np.asarray(x, dtype=np.float64)), device=device)
    return t.contiguous()


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(device=None):
    if not torch.cuda.is_available():
        raise _lib.GpfyError("gpfy_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)


def gegenbauer(level, alpha, x, with_grad=False, device=None):
    """C_level^alpha(x) (and d/dx) on the device — gegenbauer.py:63-85."""
    lib = _lib.load()
    device = require_cuda(device)
    xt = _as_dev(x, device)
    out = torch.empty_like(xt)
    dout = torch.empty_like(xt) if with_grad else None
    _lib.check(lib.gpfy_gegenbauer(_stream(device), int(level), float(alpha), _ptr(xt), xt.numel(), _ptr(out), _ptr(dout)))
    return (out, dout) if with_grad else out


def shape_function(fn, alpha, x, device=None):
    """Elementwise shape function of a Spherical kernel (`fn` from `_lib.make_shape_fn`) — spherical.py:438-455, 479-500, 589-611."""
    lib = _lib.load()
    device = require_cuda(device)
    xt = _as_dev(x, device)
    out = torch.empty_like(xt)
    _lib.check(lib.gpfy_shape_function(_stream(device), ctypes.byref(fn), float(alpha), _ptr(xt), xt.numel(), _ptr(out)))
    return out


# precision of the dense contraction C = W2 Z (include/gpfy_b200.h GPFY_PREC_*).  "f64" - the default, float64 results to the
# reference's 1e-10 contract - runs it as exact int8 digit products on tcgen05 where that kernel is built (P = 1, M <= 864, one of
# the fused backward kernels; ~1e-14 normwise) and on the FP64 tensor cores elsewhere; "f64dmma" forces the DMMA kernel everywhere.
PRECISIONS = {"f64": _lib.PREC_F64_I8, "f64i8": _lib.PREC_F64_I8, "f64dmma": _lib.PREC_F64, "f32": _lib.PREC_TF32X3, "tf32x3": _lib.PREC_TF32X3}


class HotPath:
    """One descriptor + cached scratch.  Mirrors the static configuration of
    (SphericalHarmonics, kernel, q, likelihood) in the reference."""

    def __init__(self, input_dim, m, num_outputs=1, kernel_order=1, likelihood="gaussian", q_kind="tril",
                 weight_mode="ard", device=None, projection_dim=0, precision="f64"):
        self.lib = _lib.load()
        self.device = require_cuda(device)
        self.m = [int(v) for v in m]
        self.input_dim = int(input_dim)
        self.projection_dim = int(projection_dim) if weight_mode == "projection" else 0
        self.dim = (self.projection_dim or self.input_dim) + 1
        self.M = int(sum(self.m))
        self.P = int(num_outputs)
        self.F = len(self.m)
        self.lsize = int(sum(v * v for v in self.m))
        self.scalar_w = weight_mode == "scalar"
        self.desc = _lib.make_desc(
            input_dim, self.m, num_outputs, kernel_order,
            _lib.LIK_GAUSSIAN if likelihood == "gaussian" else _lib.LIK_BERNOULLI,
            _lib.Q_TRIL if q_kind == "tril" else _lib.Q_FULLCOV,
            {"scalar": _lib.W_SCALAR, "projection": _lib.W_PROJECTION}.get(weight_mode, _lib.W_ARD), self.projection_dim,
            PRECISIONS[precision])
        self.precision = precision
        self.nw = self.input_dim * self.projection_dim if self.projection_dim else (1 if self.scalar_w else self.input_dim)
        self.layout = _lib.GpfyElboLayout()
        _lib.check(self.lib.gpfy_elbo_packed_layout(ctypes.byref(self.desc), ctypes.byref(self.layout)))
        self._ws = {}

    # ------------------------------------------------------------------
    def workspace(self, n, op):
        nbytes = ctypes.c_size_t(0)
        _lib.check(self.lib.gpfy_workspace_bytes(ctypes.byref(self.desc), int(n), int(op), ctypes.byref(nbytes)))
        key = op
        ws = self._ws.get(key)
        if ws is None or ws.numel() < nbytes.value:
            ws = torch.empty(nbytes.value, dtype=torch.uint8, device=self.device)
            self._ws[key] = ws
        return ws

    def _dev(self, x):
        return _as_dev(x, self.device)

    # ------------------------------------------------------------------
    def features(self, X, vhat, lcat, w=None, b=None, degree_scale=None, mul_r=False, apply_to_sphere=True, out=None):
        """[M, N] feature matrix and radii (spherical_harmonics.py:290-320,344-362,385-387); `out` = (Phi, r) buffers to reuse."""
        X = self._dev(X)
        n = X.shape[0]
        vhat, lcat = self._dev(vhat), self._dev(lcat)
        w = self._dev(w) if w is not None else None
        b = self._dev(b) if b is not None else None
        ds = self._dev(degree_scale) if degree_scale is not None else None
        if out is not None:
            out, r = out
            if tuple(out.shape) != (self.M, n) or tuple(r.shape) != (n,) or not (out.is_contiguous() and r.is_contiguous()):
                raise _lib.GpfyError("features: `out` must be contiguous float64 buffers of shape [M, N] and [N]")
        else:
            out = torch.empty((self.M, n), dtype=torch.float64, device=self.device)
            r = torch.empty((n,), dtype=torch.float64, device=self.device)
        ws = self.workspace(n, _lib.OP_FEATURES)
        _lib.check(self.lib.gpfy_sh_features_fwd(
            _stream(self.device), ctypes.byref(self.desc), n, _ptr(X), int(bool(apply_to_sphere)), _ptr(w), _ptr(b),
            _ptr(vhat), _ptr(lcat), _ptr(ds), int(bool(mul_r)), _ptr(out), _ptr(r), _ptr(ws), ws.numel()))
        return out, r

    def features_vjp(self, X, vhat, lcat, dPhi, w=None, b=None, degree_scale=None, mul_r=False, apply_to_sphere=True):
        """Reverse pass of `features`: cotangents {"vhat" [M, dim], "lcat" [sum m_l^2], "scale" [F], "w", "b"} of the inputs given
        the cotangent dPhi [M, N] of the feature matrix - what jax.grad builds for polynomial_expansion / Kuf / project_fun
        (spherical_harmonics.py:290-320, 344-362, 385-387)."""
        X, dPhi = self._dev(X), self._dev(dPhi)
        n = X.shape[0]
        if tuple(dPhi.shape) != (self.M, n):
            raise _lib.GpfyError("features_vjp: dPhi must be [M, N]")
        vhat, lcat = self._dev(vhat), self._dev(lcat)
        w = self._dev(w) if w is not None else None
        b = self._dev(b) if b is not None else None
        ds = self._dev(degree_scale) if degree_scale is not None else None
        z = lambda *shape: torch.zeros(shape, dtype=torch.float64, device=self.device)
        out = {"vhat": z(self.M, self.dim), "lcat": z(self.lsize), "scale": z(self.F)}
        if apply_to_sphere:
            out["w"], out["b"] = z(1 if self.scalar_w else self.input_dim), z(1)
        ws = self.workspace(n, _lib.OP_FEATURES_VJP)
        _lib.check(self.lib.gpfy_sh_features_vjp(
            _stream(self.device), ctypes.byref(self.desc), n, _ptr(X), int(bool(apply_to_sphere)), _ptr(w), _ptr(b), _ptr(vhat),
            _ptr(lcat), _ptr(ds), int(bool(mul_r)), _ptr(dPhi), _ptr(out["vhat"]), _ptr(out["lcat"]), _ptr(out["scale"]),
            _ptr(out.get("w")), _ptr(out.get("b")), _ptr(ws), ws.numel()))
        return out

    def elbo_valgrad_packed(self, X, Y, w, b, variance, eig, vhat, lcat, mu, sigma, sigma2, out=None):
        """Packed [data term, gradients] for the rows given (sums; see GpfyElboLayout)."""
        X, Y = self._dev(X), self._dev(Y)
        n = X.shape[0]
        args = [self._dev(a) for a in (w, b, variance, eig, vhat, lcat, mu, sigma)]
        s2 = self._dev(sigma2 if sigma2 is not None else 1.0)
        if out is None:
            out = torch.empty((self.layout.total,), dtype=torch.float64, device=self.device)
        ws = self.workspace(n, _lib.OP_ELBO)
        _lib.check(self.lib.gpfy_svgp_elbo_valgrad(
            _stream(self.device), ctypes.byref(self.desc), n, _ptr(X), _ptr(Y), *[_ptr(a) for a in args], _ptr(s2),
            _ptr(out), _ptr(ws), ws.numel()))
        return out

    def contraction_path(self, n, op=_lib.OP_ELBO):
        """'f64' (DMMA), 'f64i8' (int8 digit products on tcgen05, float64-accurate) or 'f32' (3 x TF32): the arithmetic the
        dense contraction C = W2 Z runs in for `n` rows on this device (gpfy_contraction_path)."""
        path = ctypes.c_int32(0)
        _lib.check(self.lib.gpfy_contraction_path(ctypes.byref(self.desc), int(max(n, 1)), int(op), ctypes.byref(path)))
        return {_lib.PREC_F64: "f64", _lib.PREC_TF32X3: "f32", _lib.PREC_F64_I8: "f64i8"}[path.value]

    # ---- host-resident data: rows are streamed through the device slab by slab ----
    def slab_rows(self, n):
        """Rows per streamed slab: one internal chunk of the kernels (gpfy_chunk_rows; 0.9 M rows at the headline shape)."""
        rows = ctypes.c_int64(0)
        _lib.check(self.lib.gpfy_chunk_rows(ctypes.byref(self.desc), int(max(n, 1)), ctypes.byref(rows)))
        return max(1, min(int(n), int(rows.value)))

    @staticmethod
    def slab_bounds(n, cap, unit=1024):
        """Row boundaries of the streamed slabs.  Nothing can overlap the copy of the FIRST slab, so it is small (two tile
        rounds, 2 * unit rows) and the slabs then grow by 1.5 x up to `cap` rows (one internal chunk): a slab's copy has to
        land while its predecessor is being processed, and with eight ranks sharing the host's memory bandwidth a copy
        costs ~0.6 of the compute time of the same rows (measured, r02c: a 5 x ramp stalled the compute stream and made
        the 8-GPU e2e figure worse than no ramp).  Every slab but the last is a multiple of `unit` = 192 * SMs rows, a
        whole number of GEMM tile rounds per SM."""
        n, cap, unit = int(n), int(cap), max(1, int(unit))
        bounds, step = [0], 2 * unit
        while bounds[-1] < n:
            step = min(cap, max(unit, step // unit * unit))
            bounds.append(min(n, bounds[-1] + step))
            step = step * 3 // 2 + unit - 1
        return bounds

    def elbo_valgrad_host(self, X_host, Y_host, w, b, variance, eig, vhat, lcat, mu, sigma, sigma2, out_host=None, reduce=None):
        """Same as `elbo_valgrad_packed` for X [N, d], Y [N, P] in (pinned) HOST memory: slabs of rows are
        copied on a side stream while the previous slab is being processed (gpfy_svgp_elbo_begin / rows /
        finish), `reduce` (the ranks' all-reduce, stream-ordered) is applied to the packed buffer on the
        device, and the result is read back into `out_host` (pinned).  Returns out_host after the device
        has finished."""
        if not (torch.is_tensor(X_host) and torch.is_tensor(Y_host)) or X_host.is_cuda or Y_host.is_cuda:
            raise _lib.GpfyError("elbo_valgrad_host expects CPU torch tensors (pin them for full copy bandwidth)")
        X_host = X_host.contiguous()
        Y_host = Y_host.reshape(X_host.shape[0], self.P).contiguous()
        n = X_host.shape[0]
        args = [self._dev(a) for a in (w, b, variance, eig, vhat, lcat, mu, sigma)]
        s2 = self._dev(sigma2 if sigma2 is not None else 1.0)
        if out_host is None:
            out_host = torch.empty((self.layout.total,), dtype=torch.float64).pin_memory()
        cap = self.slab_rows(n)
        bounds = self.slab_bounds(n, cap, 192 * torch.cuda.get_device_properties(self.device).multi_processor_count)
        ws = self.workspace(cap, _lib.OP_ELBO)
        st = self._stream_state(cap)
        dptr = ctypes.byref(self.desc)
        main = torch.cuda.current_stream(self.device)
        wv, bv, vv, ev, vh, lc, muv, sg = args
        _lib.check(self.lib.gpfy_svgp_elbo_begin(_stream(self.device), dptr, cap, _ptr(wv), _ptr(bv), _ptr(vv), _ptr(ev), _ptr(vh),
                                                 _ptr(lc), _ptr(muv), _ptr(sg), _ptr(ws), ws.numel()))
        k = 0
        for lo, hi in zip(bounds[:-1], bounds[1:]):
            xb, yb = st["X"][k % 2], st["Y"][k % 2]
            with torch.cuda.stream(st["copy"]):
                st["copy"].wait_event(st["free"][k % 2])      # slab buffer no longer read by the compute stream
                xb[:hi - lo].copy_(X_host[lo:hi], non_blocking=True)
                yb[:hi - lo].copy_(Y_host[lo:hi], non_blocking=True)
                st["ready"][k % 2].record(st["copy"])
            main.wait_event(st["ready"][k % 2])
            _lib.check(self.lib.gpfy_svgp_elbo_rows(_stream(self.device), dptr, cap, hi - lo, _ptr(xb), _ptr(yb), _ptr(vv), _ptr(s2),
                                                    _ptr(ws), ws.numel()))
            st["free"][k % 2].record(main)
            k += 1
        out = st["packed"]
        _lib.check(self.lib.gpfy_svgp_elbo_finish(_stream(self.device), dptr, cap, _ptr(wv), _ptr(bv), _ptr(vv), _ptr(ev), _ptr(muv),
                                                  _ptr(out), _ptr(ws), ws.numel()))
        if reduce is not None:
            reduce(out)
        out_host.copy_(out, non_blocking=True)
        main.synchronize()
        return out_host

    def _stream_state(self, cap):
        st = getattr(self, "_hs", None)
        if st is None or st["cap"] < cap:
            mk = lambda c: [torch.empty((cap, c), dtype=torch.float64, device=self.device) for _ in range(2)]
            st = {"cap": cap, "X": mk(self.input_dim), "Y": mk(self.P), "copy": torch.cuda.Stream(self.device),
                  "ready": [torch.cuda.Event() for _ in range(2)], "free": [torch.cuda.Event() for _ in range(2)],
                  "packed": torch.empty((self.layout.total,), dtype=torch.float64, device=self.device)}
            for e in st["free"]:
                e.record(torch.cuda.current_stream(self.device))
            self._hs = st
        return st

    def unpack(self, packed):
        """Views of a packed buffer keyed like the oracle's gradient dict."""
        L = self.layout
        M, P, F, d, dim = self.M, self.P, self.F, self.input_dim, self.dim
        g = {
            "data_term": packed[L.data_term],
            "mu": packed[L.d_mu:L.d_mu + M * P].view(M, P),
            "sigma": packed[L.d_sigma:L.d_sigma + P * M * M].view(P, M, M),
            "eig": packed[L.d_eig:L.d_eig + F],
            "vhat": packed[L.d_vhat:L.d_vhat + M * dim].view(M, dim),
            "lcat": packed[L.d_lcat:L.d_lcat + self.lsize],
            "w": packed[L.d_w:L.d_w + self.nw],
            "b": packed[L.d_b],
            "v": packed[L.d_v],
            "sigma2": packed[L.d_sigma2],
        }
        return g

    def predict_diag(self, X, w, b, variance, eig, vhat, lcat, mu, sigma):
        """(fmu, fvar), each [N, P] — model.py:201-229."""
        X = self._dev(X)
        n = X.shape[0]
        args = [self._dev(a) for a in (w, b, variance, eig, vhat, lcat, mu, sigma)]
        fmu = torch.empty((n, self.P), dtype=torch.float64, device=self.device)
        fvar = torch.empty((n, self.P), dtype=torch.float64, device=self.device)
        ws = self.workspace(n, _lib.OP_PREDICT)
        _lib.check(self.lib.gpfy_predict_diag(
            _stream(self.device), ctypes.byref(self.desc), n, _ptr(X), *[_ptr(a) for a in args], _ptr(fmu), _ptr(fvar),
            _ptr(ws), ws.numel()))
        return fmu, fvar

    def to_sphere(self, X, w, b):
        """(x_sphere [N, dim], r [N]) — spherical.py:230-249."""
        X = self._dev(X)
        n = X.shape[0]
        w, b = self._dev(w), self._dev(b)
        xs = torch.empty((n, self.dim), dtype=torch.float64, device=self.device)
        r = torch.empty((n,), dtype=torch.float64, device=self.device)
        _lib.check(self.lib.gpfy_to_sphere(_stream(self.device), ctypes.byref(self.desc), n, _ptr(X), _ptr(w), _ptr(b), _ptr(xs), _ptr(r)))
        return xs, r

    def variational_expectations(self, fmu, fvar, Y, sigma2=None):
        """Expected log-likelihood per point [N] — likelihoods.py:188-220 / :254-278."""
        fmu, fvar, Y = self._dev(fmu), self._dev(fvar), self._dev(Y)
        n = fmu.shape[0]
        s2 = self._dev(sigma2 if sigma2 is not None else 1.0)
        out = torch.empty((n,), dtype=torch.float64, device=self.device)
        _lib.check(self.lib.gpfy_variational_expectations(_stream(self.device), ctypes.byref(self.desc), n, _ptr(fmu), _ptr(fvar),
                                                          _ptr(Y), _ptr(s2), _ptr(out)))
        return out

    def project_q(self, A, mu, sigma):
        """(A^T mu [N, P], diag(A^T S A) [N, P]) for a projection A [M, N] — variational.py:288-328 / :420-441."""
        A = self._dev(A)
        n = A.shape[1]
        mu, sigma = self._dev(mu), self._dev(sigma)
        mean = torch.empty((n, self.P), dtype=torch.float64, device=self.device)
        var = torch.empty((n, self.P), dtype=torch.float64, device=self.device)
        ws = self.workspace(n, _lib.OP_PREDICT)
        _lib.check(self.lib.gpfy_project_q(_stream(self.device), ctypes.byref(self.desc), n, _ptr(A), _ptr(mu), _ptr(sigma),
                                           _ptr(mean), _ptr(var), _ptr(ws), ws.numel()))
        return mean, var

    def _kmat_workspace(self, n1, n2):
        nbytes = ctypes.c_size_t(0)
        _lib.check(self.lib.gpfy_kmat_workspace_bytes(ctypes.byref(self.desc), int(n1), int(n2), ctypes.byref(nbytes)))
        ws = self._ws.get("kmat")
        if ws is None or ws.numel() < nbytes.value:
            ws = torch.empty(max(nbytes.value, 256), dtype=torch.uint8, device=self.device)
            self._ws["kmat"] = ws
        return ws

    def kernel_matrix(self, fn, X, X2, w, b, variance, unit_diagonal=True):
        """K [N, N2] of a Spherical kernel with shape function `fn` — spherical.py:318-352 (`K`), :354-382 (`K2`,
        `unit_diagonal=False`).  X2 None: pairs of X."""
        X = self._dev(X)
        X2 = None if X2 is None else self._dev(X2)
        n1 = X.shape[0]
        n2 = n1 if X2 is None else X2.shape[0]
        w, b, variance = self._dev(w), self._dev(b), self._dev(variance)
        K = torch.empty((n1, n2), dtype=torch.float64, device=self.device)
        if n1 == 0 or n2 == 0:   # an empty X2 has a null data pointer, which the C-ABI reads as "pairs of X"
            return K
        ws = self._kmat_workspace(n1, 0 if X2 is None else n2)
        _lib.check(self.lib.gpfy_kernel_matrix(_stream(self.device), ctypes.byref(self.desc), ctypes.byref(fn), n1, _ptr(X), n2, _ptr(X2),
                                               _ptr(w), _ptr(b), _ptr(variance), int(bool(unit_diagonal)), _ptr(K), _ptr(ws), ws.numel()))
        return K

    def project_variance(self, A, sigma, subtract_identity=False, base=None):
        """base + A^T (S_p - [subtract_identity] I) A -> [P, N, N] — variational.py:330-351 / :443-464; with base = K(X) and
        subtract_identity it is the covariance of `GP.predict` (model.py:129-131, spherical_harmonics.py:391-393)."""
        A, sigma = self._dev(A), self._dev(sigma)
        base = None if base is None else self._dev(base)
        n = A.shape[1]
        out = torch.empty((self.P, n, n), dtype=torch.float64, device=self.device)
        ws = self._kmat_workspace(n, 0)
        _lib.check(self.lib.gpfy_project_variance(_stream(self.device), ctypes.byref(self.desc), n, _ptr(A), _ptr(sigma),
                                                  int(bool(subtract_identity)), _ptr(base), _ptr(out), _ptr(ws), ws.numel()))
        return out

    def inducing_basis(self, V, normalise=True):
        """(vhat [M, dim], lcat [sum m_l^2]) from the raw fundamental points — spherical_harmonics.py:194-239, 265-288."""
        V = self._dev(V)
        vhat = torch.empty((self.M, self.dim), dtype=torch.float64, device=self.device)
        lcat = torch.empty((self.lsize,), dtype=torch.float64, device=self.device)
        ws = self._basis_workspace()
        _lib.check(self.lib.gpfy_inducing_basis_fwd_ws(_stream(self.device), ctypes.byref(self.desc), _ptr(V), int(bool(normalise)),
                                                       _ptr(vhat), _ptr(lcat), _ptr(ws), ws.numel() * 8))
        return vhat, lcat

    BASIS_MAX_M = 256   # gpfy_inducing_basis_*_ws: m_l <= 64 in shared memory, up to 256 with global scratch

    def _basis_workspace(self):
        """Scratch of the inducing-basis kernels for m_l > 64 (32 * sum m_l^2 bytes; one element otherwise), cached."""
        ws = getattr(self, "_basis_ws", None)
        if ws is None:
            n = 4 * self.lsize if max(self.m) > 64 else 2
            ws = self._basis_ws = torch.empty((n,), dtype=torch.float64, device=self.device)
        return ws

    def inducing_basis_vjp(self, V, lcat, dvhat, dlcat, normalise=True):
        """dV [M, dim] from the cotangents of (vhat, lcat) — the reverse pass of `inducing_basis`."""
        V, lcat, dvhat, dlcat = self._dev(V), self._dev(lcat), self._dev(dvhat), self._dev(dlcat)
        dV = torch.empty((self.M, self.dim), dtype=torch.float64, device=self.device)
        ws = self._basis_workspace()
        _lib.check(self.lib.gpfy_inducing_basis_vjp_ws(_stream(self.device), ctypes.byref(self.desc), _ptr(V), _ptr(lcat),
                                                       int(bool(normalise)), _ptr(dvhat), _ptr(dlcat), _ptr(dV), _ptr(ws), ws.numel() * 8))
        return dV

    def prior_kl_valgrad(self, mu, sigma):
        """KL(q || N(0, I)), dKL/dmu [M, P], dKL/dSigma [P, M, M] — variational.py:225-240 with :262-286 (TriL) or :375-399
        (full covariance: Cholesky, logdet and Sigma^-1 on the device)."""
        mu, sigma = self._dev(mu), self._dev(sigma)
        M, P = self.M, self.P
        out = torch.empty((1 + M * P + P * M * M,), dtype=torch.float64, device=self.device)
        ws = self._ws.get("kl")
        need = (2 * P * M * M + P) * 8 if self.desc.q_kind == _lib.Q_FULLCOV else 0
        if need and (ws is None or ws.numel() < need):
            ws = self._ws["kl"] = torch.empty(need, dtype=torch.uint8, device=self.device)
        _lib.check(self.lib.gpfy_prior_kl_valgrad_ws(_stream(self.device), ctypes.byref(self.desc), _ptr(mu), _ptr(sigma), _ptr(out),
                                                     _ptr(ws) if need else None, need))
        return out[0], out[1:1 + M * P].view(M, P), out[1 + M * P:].view(P, M, M)


DEBUG_OPTIONS = {"chunk_mult": 0, "old_fwd": 1, "no_fused_bwd": 2, "old_features": 3, "no_bwd4": 4, "no_syrk32": 5, "syrk32_kf": 6, "no_i8": 7, "no_syrk_slice": 8, "no_features3": 9, "no_fwd3": 10}


def debug_option(name, value):
    """Development switch of the library (tests / tools): `chunk_mult` = rows per internal chunk as a multiple of
    192 * SMs (0 = default); `no_fused_bwd` = 1 sends the headline shapes through the general fused backward kernel
    (zonal_bwd4), `no_bwd4` = 1 on top of it through the legacy one-point-per-lane kernels; `old_fwd`, `old_features` = 1
    select the legacy forward / two-kernel feature path."""
    _lib.check(_lib.load().gpfy_debug_set_option(DEBUG_OPTIONS[name], int(value)))


def launch_count():
    return int(_lib.load().gpfy_debug_launch_count())


KERNEL_CLASSES = ("zonal_fwd", "gemm_dmma", "lik_point", "zonal_bwd", "syrk_dmma", "dvhat_dmma", "i8_slice")


def kernel_timing(enable):
    """Bracket the per-chunk kernels of the ELBO step with CUDA events on the launching stream."""
    _lib.check(_lib.load().gpfy_debug_kernel_timing(int(bool(enable))))


def kernel_times():
    """{kernel class: (total ms, launches)} since kernel_timing(True); waits for the recorded events."""
    lib = _lib.load()
    out = {}
    for i, name in enumerate(KERNEL_CLASSES):
        ms, n = ctypes.c_double(0.0), ctypes.c_int64(0)
        _lib.check(lib.gpfy_debug_kernel_time(i, ctypes.byref(ms), ctypes.byref(n)))
        out[name] = (ms.value, n.value)
    return out
