"""CPU float64 restatement of GPfY's hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` leg may
import this module.  Nothing under `gpfy_b200/` does.

Every function restates, in `torch.float64` on the CPU, the arithmetic of the reference lines it
cites (paths relative to /root/reference/src/gpfy).  torch is used instead of numpy so that
`torch.autograd` supplies the reference gradients (the stand-in for `jax.value_and_grad`,
optimization.py:117); the Gegenbauer recurrence carries the reference's custom VJP
(gegenbauer.py:70-81) through `_GegenbauerFn`.

Parity pin: JAX cannot be installed in this image, so the reference cannot run on real JAX/XLA.
Instead the *unmodified reference source files* are executed on the numpy stand-in for the JAX API in
`oracle/jax_shim/` by `tests/golden/make_golden.py`, and this module is checked against those golden
vectors (tests/test_oracle_golden.py) plus every known-answer/invariant test the reference's own
tests hold for this path (tests/test_oracle_kats.py).  Gradients are pinned by central finite
differences of the reference forward (same golden files).  Third-party arithmetic that the shim
restates rather than executes: XLA:CPU kernels (IEEE fp64 dot/cholesky/triangular_solve) and TFP
(`FillTriangular`, `Normal/Bernoulli.log_prob`, `interp_regular_1d_grid`), all unpinned in the
reference's requirements.txt.

Parameter pack ("hot-path pack", constrained space) used by the functions below:
    dim        ambient dimension = sphere_dim + 1        (spherical_harmonics.py:99-100)
    alpha      (dim - 2) / 2                             (spherical_harmonics.py:149)
    V          list over degrees of [m_l, dim] tensors   (params["variational"]["inducing_features"])
    L          list of [m_l, m_l] lower Cholesky factors, or None => learned-V mode (NaN sentinel,
               spherical_harmonics.py:209-217,233-239): rows of V are normalised and L recomputed.
    w          weight_variances: [d] (ARD), [] (scalar) or [d, p] (linear projection)
    b          bias_variance, v = kernel variance, eig = per-degree eigenvalues lambda_l [F]
    order      kernel order (1 for NTK / PolynomialDecay)
    mu [M,P], Lq [P,M,M] (TriL) or Sigma [P,M,M] (full covariance), sigma2 (Gaussian noise)
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
from scipy.special import comb, loggamma, roots_legendre

DT = torch.float64


def T(x) -> torch.Tensor:
    return torch.as_tensor(np.asarray(x, dtype=np.float64) if not torch.is_tensor(x) else x, dtype=DT)


# ----------------------------------------------------------------------------------------------
# harmonic counting  (spherical_harmonics.py:31-47)
# ----------------------------------------------------------------------------------------------
def num_harmonics(dim: int, frequency: int) -> int:
    if frequency == 0:
        return 1
    c1 = comb(frequency + dim - 3, frequency - 1)
    c2 = comb(frequency + dim - 2, frequency)
    return int(c1 + c2)


# ----------------------------------------------------------------------------------------------
# Gegenbauer recurrence + custom VJP  (gegenbauer.py:27-85)
# ----------------------------------------------------------------------------------------------
def _gegenbauer(level: int, alpha: float, x: torch.Tensor) -> torch.Tensor:
    """C_0 = 1, C_1 = 2 a x, C_n = (2 x (n + a - 1) C_{n-1} - (n + 2a - 2) C_{n-2}) / n."""
    level = int(level)
    if level == 0:  # gegenbauer.py:56-58
        return torch.ones_like(x)
    C_prev = torch.ones_like(x)
    C = 2 * alpha * x
    n = 2.0
    while n <= level:  # gegenbauer.py:47-54; same association order as :49
        C, C_prev = (2 * x * (n + alpha - 1) * C - (n + 2 * alpha - 2) * C_prev) / n, C
        n += 1.0
    return C


class _GegenbauerFn(torch.autograd.Function):
    """d/dx C_n^a(x) = 2 a C_{n-1}^{a+1}(x), zero at level 0  (gegenbauer.py:70-81)."""

    @staticmethod
    def forward(ctx, x, level, alpha):
        with torch.no_grad():
            geg = _gegenbauer(level, alpha, x)
            grad = torch.zeros_like(x) if level == 0 else 2 * alpha * _gegenbauer(level - 1, alpha + 1, x)
        ctx.save_for_backward(grad)
        return geg

    @staticmethod
    def backward(ctx, cot):
        (grad,) = ctx.saved_tensors
        return grad * cot, None, None


def gegenbauer(level: int, alpha: float, x: torch.Tensor) -> torch.Tensor:
    return _GegenbauerFn.apply(x, int(level), float(alpha))


def gegenbauer_grad(level: int, alpha: float, x: torch.Tensor) -> torch.Tensor:
    if int(level) == 0:
        return torch.zeros_like(x)
    return 2 * alpha * _gegenbauer(int(level) - 1, alpha + 1, x)


# ----------------------------------------------------------------------------------------------
# kernel side  (spherical.py)
# ----------------------------------------------------------------------------------------------
def scale_X(w: torch.Tensor, X: torch.Tensor) -> torch.Tensor:
    """spherical.py:214-228: X @ W for a 2-D weight, else X * sqrt(w)."""
    if w.dim() == 2:
        return X @ w
    return X * torch.sqrt(w)


def to_sphere(w, b, X) -> Tuple[torch.Tensor, torch.Tensor]:
    """spherical.py:230-249: always append sqrt(bias_variance), then normalise."""
    scaled = scale_X(w, X)
    bias = torch.ones(scaled.shape[:-1] + (1,), dtype=DT) * torch.sqrt(b)
    Xb = torch.cat([scaled, bias], dim=-1)
    r = torch.sqrt(torch.sum(torch.square(Xb), dim=-1))
    return Xb / r[..., None], r


def k_diag(v, r, order: int) -> torch.Tensor:
    """spherical.py:384-400: variance * r ** (2 * order)."""
    return v * r ** (2 * order)


def polydecay_const_factor(alpha: float, truncation_level: int) -> torch.Tensor:
    """spherical.py:526-561: alpha / (n + alpha) / C_n^alpha(1), n < truncation_level.

    The reference reads C_n(1) from the lookup table; interpolating at x = 1 returns the last grid
    node, which is the recurrence evaluated at exactly 1.0 (gegenbauer.py:113-119,167)."""
    one = torch.ones((), dtype=DT)
    C1 = torch.stack([_gegenbauer(n, alpha, one) for n in range(truncation_level)])
    n = torch.arange(truncation_level, dtype=DT)
    return alpha / (n + alpha) / C1


def polydecay_eigenvalues(beta, alpha: float, truncation_level: int, levels: Sequence[int]) -> torch.Tensor:
    """spherical.py:563-587 incl. the clip of levels to truncation_level - 1 (:586)."""
    n = torch.arange(truncation_level, dtype=DT)
    decay = (1 + n) ** (-beta)
    eigval = decay * polydecay_const_factor(alpha, truncation_level) / torch.sum(decay)
    lv = torch.clamp(torch.as_tensor(list(levels), dtype=torch.long), 0, truncation_level - 1)
    return eigval[lv]


def _kappa(u: torch.Tensor, order: int) -> torch.Tensor:
    """spherical.py:279-316."""
    if order == 0:
        return (math.pi - torch.arccos(u)) / math.pi
    if order == 1:
        return (u * (math.pi - torch.arccos(u)) + torch.sqrt(1.0 - torch.square(u))) / math.pi
    return ((1.0 + 2.0 * torch.square(u)) / 3 * (math.pi - torch.arccos(u)) + torch.sqrt(1.0 - torch.square(u))) / math.pi


def ntk_shape_function(x: torch.Tensor, depth: int) -> torch.Tensor:
    """spherical.py:479-500."""
    x = torch.clamp(x, -1.0, 1.0)
    a, y = x, x
    for _ in range(depth):
        a, y = _kappa(a, 1), y * _kappa(a, 0) + _kappa(a, 1)
    return y / (depth + 1)


def arccosine_shape_function(x: torch.Tensor, depth: int, order: int) -> torch.Tensor:
    """spherical.py:438-455."""
    y = torch.clamp(x, -1.0, 1.0)
    for _ in range(depth):
        y = _kappa(y, order)
    return y


def gegenbauer_from_lookup(level: int, alpha: float, x: torch.Tensor, grid_resolution: int = 100_000) -> torch.Tensor:
    """gegenbauer.py:111-119,152-167: linear interpolation of C_level^alpha on linspace(-1, 1, grid_resolution)
    (tfp.math.interp_regular_1d_grid: fractional index (x - lo) / (hi - lo) * (G - 1), clipped)."""
    grid = _gegenbauer(level, alpha, torch.linspace(-1, 1, grid_resolution, dtype=DT))
    xi = torch.clamp((x + 1.0) / 2.0 * (grid_resolution - 1), 0.0, grid_resolution - 1.0)
    lo = torch.floor(xi)
    hi = torch.clamp(lo + 1, max=grid_resolution - 1)
    t = xi - lo
    return (1 - t) * grid[lo.long()] + t * grid[hi.long()]


def polydecay_shape_function(beta, alpha: float, truncation_level: int, x: torch.Tensor) -> torch.Tensor:
    """spherical.py:589-611: sum_n eig_n (n + alpha) / alpha C_n^alpha(x), C from the lookup table."""
    x = torch.clamp(x, -1.0, 1.0)
    levels = list(range(truncation_level))
    eig = polydecay_eigenvalues(beta, alpha, truncation_level, levels) * (torch.arange(truncation_level, dtype=DT) + alpha) / alpha
    return sum(eig[n] * gegenbauer_from_lookup(n, alpha, x) for n in levels)


def shape_function(pack, x: torch.Tensor) -> torch.Tensor:
    """Dispatch on the kernel class of the pack (spherical.py:438-455, 479-500, 589-611)."""
    kind = pack.get("kernel", "NTK")
    if kind == "PolynomialDecay":
        return polydecay_shape_function(pack["beta"], pack["alpha"], pack["truncation_level"], x)
    if kind == "NTK":
        return ntk_shape_function(x, pack["depth"])
    return arccosine_shape_function(x, pack["depth"], pack["order"])


def K(pack, X: torch.Tensor, X2: Optional[torch.Tensor] = None, unit_diagonal: bool = True) -> torch.Tensor:
    """spherical.py:318-352 (`K`: the diagonal cosine is set to exactly 1 when X2 is None) and
    :354-382 (`K2`: `unit_diagonal=False`, no such fix-up; X2 defaults to X)."""
    xs, r1 = to_sphere(pack["w"], pack["b"], X)
    if X2 is None:
        xs2, r2 = xs, r1
        c = xs @ xs2.T
        if unit_diagonal:
            c = c.clone()
            c.fill_diagonal_(1.0)
    else:
        xs2, r2 = to_sphere(pack["w"], pack["b"], X2)
        c = xs @ xs2.T
    k = shape_function(pack, c)
    return pack["v"] * k * (r1[:, None] * r2[None, :]) ** pack["order"]


_LEG_X, _LEG_W = roots_legendre(1000)  # harmonics/utils.py:23-25


def funk_hecke_lambda(fn: Callable[[torch.Tensor], torch.Tensor], n: int, d: int) -> torch.Tensor:
    """harmonics/utils.py:42-71: 1000-node Gauss-Legendre Funk-Hecke eigenvalue."""
    alpha = (d - 2.0) / 2.0
    C1 = _gegenbauer(n, alpha, torch.ones((), dtype=DT))
    ratio = math.exp(loggamma(d / 2) - loggamma((d - 1) / 2)) / math.sqrt(math.pi)
    x = T(_LEG_X)
    integrand = _gegenbauer(n, alpha, x) * fn(x) * (1.0 - torch.square(x)) ** ((d - 3) / 2)
    return ratio / C1 * torch.dot(T(_LEG_W), integrand)


def spherical_eigenvalues(fn, sphere_dim: int, max_num_eigvals: int = 20) -> torch.Tensor:
    """spherical.py:153-180: eigenvalues for degrees 0..max_num_eigvals-1 (NTK / ArcCosine)."""
    return torch.stack([funk_hecke_lambda(fn, n, sphere_dim + 1) for n in range(max_num_eigvals)])


# ----------------------------------------------------------------------------------------------
# spherical-harmonic features  (spherical_harmonics.py)
# ----------------------------------------------------------------------------------------------
def normalise_V(V: List[torch.Tensor]) -> List[torch.Tensor]:
    """spherical_harmonics.py:211-215."""
    return [v / torch.sqrt(torch.sum(torch.square(v), dim=1, keepdim=True)) for v in V]


def orthogonalise_basis(Vhat: List[torch.Tensor], alpha: float) -> List[torch.Tensor]:
    """spherical_harmonics.py:265-288: L_l = chol(alpha/(alpha+l) C_l(V V^T) + 1e-16 I)."""
    out = []
    for n, v in enumerate(Vhat):
        c = alpha / (alpha + float(n))
        B = c * gegenbauer(n, alpha, v @ v.T)
        out.append(torch.linalg.cholesky(B + 1e-16 * torch.eye(B.shape[0], dtype=DT)))
    return out


def resolve_V_L(pack) -> Tuple[List[torch.Tensor], List[torch.Tensor]]:
    """NaN-sentinel rule (spherical_harmonics.py:194-239): L is None <=> learned-V mode."""
    if pack.get("L") is None:
        Vhat = normalise_V(pack["V"])
        return Vhat, orthogonalise_basis(Vhat, pack["alpha"])
    return list(pack["V"]), list(pack["L"])


def polynomial_expansion(Vhat, L, alpha: float, Xs: torch.Tensor) -> torch.Tensor:
    """spherical_harmonics.py:290-320: Phi = concat_l L_l^{-1} C_l^alpha(V_l Xs^T)  -> [M, N]."""
    rows = []
    for n, (v, Ln) in enumerate(zip(Vhat, L)):
        zonal = gegenbauer(n, alpha, v @ Xs.T)
        rows.append(torch.linalg.solve_triangular(Ln, zonal, upper=False))
    return torch.cat(rows, dim=0)


def Kuu(eig: torch.Tensor, m_per_degree: Sequence[int]) -> torch.Tensor:
    """spherical_harmonics.py:322-342: 1/lambda_l repeated m_l times."""
    return torch.cat([torch.ones(m, dtype=DT) / e for e, m in zip(eig, m_per_degree)])


def Kuf(pack, X: torch.Tensor) -> torch.Tensor:
    """spherical_harmonics.py:344-362: Phi * r * sqrt(variance) (r to the power 1, whatever the order)."""
    Vhat, L = resolve_V_L(pack)
    Xs, r = to_sphere(pack["w"], pack["b"], X)
    sh = polynomial_expansion(Vhat, L, pack["alpha"], Xs)
    return sh * r[None, :] * torch.sqrt(pack["v"])


def projection(pack, X: torch.Tensor) -> torch.Tensor:
    """spherical_harmonics.py:385-387: A = sqrt(1/Kuu)[:, None] * Kuf."""
    m = [v.shape[0] for v in pack["V"]]
    return torch.sqrt(1 / Kuu(pack["eig"], m))[:, None] * Kuf(pack, X)


# ----------------------------------------------------------------------------------------------
# variational distribution  (variational.py)
# ----------------------------------------------------------------------------------------------
def project_mean(mu: torch.Tensor, A: torch.Tensor) -> torch.Tensor:
    """variational.py:288-305: A^T mu -> [N, P]."""
    return A.T @ mu


def project_diag_variance_tril(Lq: torch.Tensor, A: torch.Tensor) -> torch.Tensor:
    """variational.py:307-328: sum_m (Lq^T A)^2 -> [N, P]."""
    tmp = Lq.transpose(-1, -2) @ A[None]
    return torch.sum(torch.square(tmp), dim=-2).T


def project_diag_variance_full(S: torch.Tensor, A: torch.Tensor) -> torch.Tensor:
    """variational.py:420-441: sum_m (S A) * A -> [N, P]."""
    return torch.sum((S @ A[None]) * A[None], dim=-2).T


def project_variance_tril(Lq: torch.Tensor, A: torch.Tensor) -> torch.Tensor:
    """variational.py:330-351: (Lq^T A)^T (Lq^T A) -> [P, N, N]."""
    tmp = Lq.transpose(-1, -2) @ A[None]
    return tmp.transpose(-1, -2) @ tmp


def project_variance_full(S: torch.Tensor, A: torch.Tensor) -> torch.Tensor:
    """variational.py:443-464: A^T S A -> [P, N, N]."""
    return A[None].transpose(-1, -2) @ (S @ A[None])


def prior_KL(mu: torch.Tensor, cov: torch.Tensor, tril: bool = True) -> torch.Tensor:
    """variational.py:225-240 with logdet/trace of :262-286 (TriL) or :375-399 (full)."""
    KL = -0.5 * torch.tensor(float(mu.numel()), dtype=DT)
    if tril:
        logdet = torch.sum(torch.log(torch.square(torch.diagonal(cov, dim1=-2, dim2=-1))), dim=-1)
        trace = torch.sum(torch.square(cov), dim=(-1, -2))
    else:
        Lc = torch.linalg.cholesky(cov)
        logdet = torch.sum(torch.log(torch.square(torch.diagonal(Lc, dim1=-2, dim2=-1))), dim=-1)
        trace = torch.diagonal(cov, dim1=-2, dim2=-1).sum(-1)
    KL = KL - 0.5 * torch.sum(logdet)
    KL = KL + 0.5 * torch.sum(torch.square(mu))
    KL = KL + 0.5 * torch.sum(trace)
    return KL


# ----------------------------------------------------------------------------------------------
# likelihoods  (likelihoods.py, quadrature.py)
# ----------------------------------------------------------------------------------------------
def gaussian_var_exp(sigma2, Fmu, Fvar, Y) -> torch.Tensor:
    """likelihoods.py:188-220."""
    return torch.sum(-0.5 * math.log(2 * math.pi) - 0.5 * torch.log(sigma2) - 0.5 * ((Y - Fmu) ** 2 + Fvar) / sigma2, dim=-1)


GH_X, GH_W = np.polynomial.hermite.hermgauss(20)  # quadrature.py:26-27


def inv_probit(x: torch.Tensor) -> torch.Tensor:
    """likelihoods.py:281-292."""
    jitter = 1e-3
    return 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0))) * (1 - 2 * jitter) + jitter


def bernoulli_log_prob(F: torch.Tensor, Y: torch.Tensor) -> torch.Tensor:
    """likelihoods.py:65-79,236-252: sum over outputs of TFP Bernoulli(probs=inv_probit(F)).log_prob(Y),
    i.e. (1-y) log1p(-p) + y log p."""
    p = inv_probit(F)
    return torch.sum((1 - Y) * torch.log1p(-p) + Y * torch.log(p), dim=-1)


def bernoulli_var_exp(Fmu, Fvar, Y) -> torch.Tensor:
    """likelihoods.py:254-278 + quadrature.py:30-53: 20-node Gauss-Hermite."""
    Xq = Fmu[..., None, :] + math.sqrt(2.0) * torch.sqrt(Fvar)[..., None, :] * T(GH_X)[:, None]
    W = T(GH_W) / math.sqrt(math.pi)
    return torch.sum(bernoulli_log_prob(Xq, Y[..., None, :]) * W, dim=-1)


# ----------------------------------------------------------------------------------------------
# model + objective  (model.py:201-229, optimization.py:125-156)
# ----------------------------------------------------------------------------------------------
def predict_diag(pack, X: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """model.py:220-225 with conditional() of :103-145."""
    A = projection(pack, X)
    fmu = project_mean(pack["mu"], A)
    _, r = to_sphere(pack["w"], pack["b"], X)
    cond_var = k_diag(pack["v"], r, pack["order"]) - torch.sum(torch.square(A), dim=-2)  # spherical_harmonics.py:388-390
    if "Lq" in pack:
        qv = project_diag_variance_tril(pack["Lq"], A)
    else:
        qv = project_diag_variance_full(pack["Sigma"], A)
    return fmu, cond_var[..., None] + qv  # model.py:136-138


def predict(pack, X: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """model.py:231-259 with conditional() of :103-145: mean [N, P], covariance [P, N, N]
    = K(X) - A^T A (spherical_harmonics.py:391-393) + the projected q covariance."""
    A = projection(pack, X)
    fmu = project_mean(pack["mu"], A)
    cond_cov = K(pack, X) - A.T @ A
    if "Lq" in pack:
        qc = project_variance_tril(pack["Lq"], A)
    else:
        qc = project_variance_full(pack["Sigma"], A)
    return fmu, cond_cov + qc


def var_exp(pack, fmu, fvar, Y) -> torch.Tensor:
    if pack["lik"] == "gaussian":
        return gaussian_var_exp(pack["sigma2"], fmu, fvar, Y)
    return bernoulli_var_exp(fmu, fvar, Y)


def elbo(pack, X: torch.Tensor, Y: torch.Tensor, dataset_size: int = -1) -> torch.Tensor:
    """optimization.py:151-156."""
    fmu, fvar = predict_diag(pack, X)
    ve = var_exp(pack, fmu, fvar, Y)
    scale = dataset_size if dataset_size > 0 else X.shape[0]
    if "Lq" in pack:
        KL = prior_KL(pack["mu"], pack["Lq"], tril=True)
    else:
        KL = prior_KL(pack["mu"], pack["Sigma"], tril=False)
    return scale * torch.mean(ve, dim=-1) - KL


def data_term_sum(pack, X, Y) -> torch.Tensor:
    """sum_n var_exp_n: the N-dependent (shardable) part of the ELBO."""
    fmu, fvar = predict_diag(pack, X)
    return torch.sum(var_exp(pack, fmu, fvar, Y))


# ----------------------------------------------------------------------------------------------
# helpers for tests / bench
# ----------------------------------------------------------------------------------------------
GRAD_KEYS = ("mu", "Lq", "Sigma", "sigma2", "w", "b", "v", "eig")


def pack_to_torch(pack: Dict, requires_grad: bool = False, grad_V: bool = True, grad_L: bool = True) -> Dict:
    """Copies a numpy pack into float64 torch leaves (optionally requiring grad)."""
    out = dict(pack)
    for k in GRAD_KEYS:
        if k in pack and pack[k] is not None:
            out[k] = T(pack[k]).clone().requires_grad_(requires_grad)
    out["V"] = [T(v).clone().requires_grad_(requires_grad and grad_V) for v in pack["V"]]
    if pack.get("L") is not None:
        out["L"] = [T(l).clone().requires_grad_(requires_grad and grad_L) for l in pack["L"]]
    return out


def elbo_and_grads(pack: Dict, X, Y, dataset_size: int = -1, data_only: bool = False):
    """ELBO (or the bare data-term sum) and autograd gradients w.r.t. every constrained leaf,
    V and L treated as independent inputs when pack["L"] is given (the C-ABI's contract)."""
    tp = pack_to_torch(pack, requires_grad=True)
    Xt, Yt = T(X), T(Y)
    val = data_term_sum(tp, Xt, Yt) if data_only else elbo(tp, Xt, Yt, dataset_size)
    leaves, names = [], []
    for k in GRAD_KEYS:
        if k in tp and torch.is_tensor(tp[k]) and tp[k].requires_grad:
            leaves.append(tp[k]); names.append(k)
    for i, vv in enumerate(tp["V"]):
        leaves.append(vv); names.append(f"V{i}")
    if tp.get("L") is not None:
        for i, ll in enumerate(tp["L"]):
            leaves.append(ll); names.append(f"L{i}")
    grads = torch.autograd.grad(val, leaves, allow_unused=True)
    g = {n: (torch.zeros_like(l) if gr is None else gr).detach().numpy() for n, l, gr in zip(names, leaves, grads)}
    return float(val.detach()), g
