"""numpy blueprint of the algorithm the CUDA path implements (design check, not product code).

Folded formulation: with E = diag(sqrt(eig_l * v)) * blockdiag(L_l^{-1}) the projection is
a_n = r_n E z_n, where z_n = C_l^alpha(Vhat x_n) are the *un-whitened* zonal values.  Then

    fmu_np  = r_n (E^T mu_p) . z_n                      = r_n mu2_p . z_n
    fvar_np = v r_n^{2 order} + r_n^2 z_n^T (E^T W0_p E) z_n = v r^{2o} + r^2 z^T W2_p z
    W0_p    = Lq_p Lq_p^T - I   (TriL)   |   Sigma_p - I   (full covariance)

so the N-dependent work is one symmetric GEMM  C = W2 Z, one weighted SYRK  Gz = Z diag(q r^2) Z^T,
thin reductions, and the zonal forward/backward; every L_l^{-1} moves into O(M^3) per-step matrix
algebra.  Run:  python tests/folded_prototype.py   (compares with oracle autograd).
"""
import math
import os
import sys

import numpy as np
from scipy.linalg import solve_triangular

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def geg_and_deriv(level, alpha, t):
    """C_level^alpha(t) and d/dt = 2 alpha C_{level-1}^{alpha+1}(t)."""
    def rec(n, a):
        if n == 0:
            return np.ones_like(t)
        Cp, C = np.ones_like(t), 2 * a * t
        for k in range(2, n + 1):
            C, Cp = (2 * t * (k + a - 1) * C - (k + 2 * a - 2) * Cp) / k, C
        return C
    z = rec(level, alpha)
    dz = np.zeros_like(t) if level == 0 else 2 * alpha * rec(level - 1, alpha + 1)
    return z, dz


GH_X, GH_W = np.polynomial.hermite.hermgauss(20)


def lik_terms(lik, fmu, fvar, Y, sigma2):
    """e_n, p = de/dfmu, q = de/dfvar  (per point, per output)."""
    if lik == "gaussian":
        e = -0.5 * math.log(2 * math.pi) - 0.5 * np.log(sigma2) - 0.5 * ((Y - fmu) ** 2 + fvar) / sigma2
        p = (Y - fmu) / sigma2
        q = np.full_like(fmu, -0.5 / sigma2)
        ds2 = np.sum(-0.5 / sigma2 + 0.5 * ((Y - fmu) ** 2 + fvar) / sigma2 ** 2)
        return e, p, q, ds2
    from scipy.special import erf
    sd = np.sqrt(fvar)
    e = np.zeros_like(fmu); p = np.zeros_like(fmu); q = np.zeros_like(fmu)
    for xi, wi in zip(GH_X, GH_W / math.sqrt(math.pi)):
        f = fmu + math.sqrt(2.0) * sd * xi
        pr = 0.5 * (1 + erf(f / math.sqrt(2.0))) * (1 - 2e-3) + 1e-3
        e += wi * ((1 - Y) * np.log1p(-pr) + Y * np.log(pr))
        dl = (Y / pr - (1 - Y) / (1 - pr)) * (1 - 2e-3) * np.exp(-0.5 * f * f) / math.sqrt(2 * math.pi)
        p += wi * dl
        q += wi * dl * xi / (math.sqrt(2.0) * sd)
    return e, p, q, 0.0


def folded_elbo_data_and_grads(pack, X, Y):
    """Returns sum_n var_exp_n and gradients w.r.t. the constrained C-ABI inputs."""
    V, L = pack["V"], pack["L"]
    F = len(V)
    m = [v.shape[0] for v in V]
    off = np.concatenate([[0], np.cumsum(m)])
    M = off[-1]
    alpha, order = pack["alpha"], pack["order"]
    w, b, v, eig = pack["w"], float(pack["b"]), float(pack["v"]), pack["eig"]
    mu = pack["mu"]
    P = mu.shape[1]
    tril = "Lq" in pack
    cov = pack["Lq"] if tril else pack["Sigma"]

    # ---- prologue: O(M^3) ----
    dvec = np.concatenate([np.full(m[l], math.sqrt(eig[l] * v)) for l in range(F)])
    Linv = np.zeros((M, M))
    for l in range(F):
        Linv[off[l]:off[l + 1], off[l]:off[l + 1]] = solve_triangular(L[l], np.eye(m[l]), lower=True)
    E = dvec[:, None] * Linv
    W0 = [(cov[p] @ cov[p].T if tril else cov[p]) - np.eye(M) for p in range(P)]
    W2 = [E.T @ W0[p] @ E for p in range(P)]
    mu2 = [E.T @ mu[:, p] for p in range(P)]

    # ---- zonal forward ----
    N = X.shape[0]
    u = np.concatenate([X * np.sqrt(w), np.full((N, 1), math.sqrt(b))], 1)
    r = np.sqrt(np.sum(u * u, 1))
    xh = u / r[:, None]
    Z = np.zeros((M, N)); dZdt = np.zeros((M, N))
    for l in range(F):
        t = V[l] @ xh.T
        Z[off[l]:off[l + 1]], dZdt[off[l]:off[l + 1]] = geg_and_deriv(l, alpha, t)

    # ---- GEMM + per-point likelihood ----
    fmu = np.stack([r * (mu2[p] @ Z) for p in range(P)], 1)
    C = [W2[p] @ Z for p in range(P)]
    psi = np.stack([np.sum(Z * C[p], 0) for p in range(P)], 1)
    fvar = v * r[:, None] ** (2 * order) + r[:, None] ** 2 * psi
    e, pp, qq, ds2 = lik_terms(pack["lik"], fmu, fvar, Y, float(pack.get("sigma2", 1.0)))
    val = float(np.sum(e))

    # ---- weighted SYRK + thin reductions ----
    Gz = [(Z * (qq[:, p] * r ** 2)[None]) @ Z.T for p in range(P)]
    bz = [Z @ (pp[:, p] * r) for p in range(P)]
    dv_direct = float(np.sum(qq * r[:, None] ** (2 * order)))

    # ---- zonal backward ----
    dZ = sum(mu2[p][:, None] * (r * pp[:, p])[None] + 2 * C[p] * (qq[:, p] * r ** 2)[None] for p in range(P))
    dr = sum(pp[:, p] * (mu2[p] @ Z) + qq[:, p] * (2 * order * v * r ** (2 * order - 1) + 2 * r * psi[:, p]) for p in range(P))
    dT = dZ * dZdt
    dV = [dT[off[l]:off[l + 1]] @ xh for l in range(F)]
    Vcat = np.concatenate(V, 0)
    dxh = dT.T @ Vcat
    du = (dxh - xh * np.sum(xh * dxh, 1, keepdims=True)) / r[:, None] + dr[:, None] * xh
    dw = np.sum(du[:, :-1] * X, 0) / (2 * np.sqrt(w))
    db = float(np.sum(du[:, -1]) / (2 * math.sqrt(b)))

    # ---- epilogue: O(M^3) ----
    g = {"sigma2": ds2, "w": dw, "b": db}
    g["mu"] = np.stack([E @ bz[p] for p in range(P)], 1)
    dW0 = [E @ Gz[p] @ E.T for p in range(P)]
    if tril:
        g["Lq"] = np.stack([2 * dW0[p] @ cov[p] for p in range(P)])
    else:
        g["Sigma"] = np.stack(dW0)
    dE = sum(np.outer(mu[:, p], bz[p]) + 2 * W0[p] @ E @ Gz[p] for p in range(P))
    dd = np.sum(dE * Linv, 1)
    g["eig"] = np.array([np.sum(dd[off[l]:off[l + 1]] * dvec[off[l]:off[l + 1]]) / (2 * eig[l]) for l in range(F)])
    g["v"] = dv_direct + float(np.sum(dd * dvec) / (2 * v))
    dLinv = dvec[:, None] * dE
    for l in range(F):
        blk = np.tril(dLinv[off[l]:off[l + 1], off[l]:off[l + 1]])  # Linv is lower-triangular
        Li = Linv[off[l]:off[l + 1], off[l]:off[l + 1]]
        g[f"L{l}"] = -(Li.T @ blk @ Li.T)
        g[f"V{l}"] = dV[l]
    return val, g


if __name__ == "__main__":
    from oracle import gpfy_oracle as O
    from tests.golden_util import CASES, golden_pack, load, rel_err
    for name in CASES:
        gd = load(name)
        pack = golden_pack(gd)
        if pack["w"].ndim == 2:
            continue  # projection weights: handled by a host-side pre-multiply
        tp = O.pack_to_torch(pack)
        Vh, L = O.resolve_V_L(tp)
        pack["V"] = [x.numpy() for x in Vh]
        pack["L"] = [x.numpy() for x in L]
        val, g = folded_elbo_data_and_grads(pack, gd["X"], gd["Y"])
        oval, og = O.elbo_and_grads(pack, gd["X"], gd["Y"], data_only=True)
        worst = 0.0
        for k, ref in og.items():
            mine = g[k]
            if k.startswith("L") and k != "Lq":
                mine, ref = np.tril(mine), np.tril(ref)
            worst = max(worst, rel_err(mine, ref) if np.max(np.abs(ref)) > 0 else float(np.max(np.abs(mine))))
        print(f"{name:28s} value rel {abs(val - oval) / abs(oval):.2e}  worst grad rel {worst:.2e}")
