"""Seeded synthetic problems of BASELINE.json's shapes: one global data set per configuration, the same for every world size.

Used by bench.py, smoke() and the full-size GPU tests.  There is no network for datasets; the data
are N(0, I) inputs with a smooth target, the parameters a random but well-conditioned point.
"""
import math

import numpy as np
import torch

from .ops import HotPath


def num_harmonics(dim, n):
    """N(dim, n) of spherical_harmonics.py:31-47 (exact integer arithmetic)."""
    if n == 0:
        return 1
    return math.comb(n + dim - 3, n - 1) + math.comb(n + dim - 2, n)


def algorithmic_flops_per_point(dim, m, with_grad=True):
    """proj + geg + trsv (+ M^2 + 4M forward; + 2M^2 + trsv + 2 geg + 2 proj backward)."""
    M = sum(m)
    proj = 2 * dim * M
    geg = sum(ml * (4 * max(l - 1, 0) + (1 if l >= 1 else 0)) for l, ml in enumerate(m))
    trsv = sum(ml * ml for ml in m)
    phi = proj + geg + trsv
    if not with_grad:
        return phi
    return (phi + M * M + 4 * M) + (2 * M * M + trsv + 2 * geg + 2 * proj)


def _gegenbauer_np(level, alpha, x):
    if level == 0:
        return np.ones_like(x)
    cp, c = np.ones_like(x), 2 * alpha * x
    for k in range(2, level + 1):
        c, cp = (2 * x * (k + alpha - 1) * c - (k + 2 * alpha - 2) * cp) / k, c
    return c


def synthetic_basis(d, F, T, rng):
    """Vhat rows (unit norm) and Cholesky factors per degree: degree 0 -> e_1, degree 1 -> identity
    (valid fundamental systems), truncated degrees -> random unit vectors (spherical_harmonics.py:123-127)."""
    dim = d + 1
    alpha = (dim - 2.0) / 2.0
    V, L, m = [], [], []
    for l in range(F):
        nl = num_harmonics(dim, l)
        ml = min(T, nl)
        if l == 0:
            v = np.zeros((1, dim)); v[0, 0] = 1.0
        elif l == 1 and ml == dim:
            v = np.eye(dim)
        else:
            v = rng.standard_normal((ml, dim))
            v /= np.linalg.norm(v, axis=1, keepdims=True)
        B = alpha / (alpha + l) * _gegenbauer_np(l, alpha, v @ v.T)
        L.append(np.linalg.cholesky(B + 1e-16 * np.eye(ml)))
        V.append(v); m.append(ml)
    return V, L, m


def headline_params(d=8, F=11, T=30, P=1, likelihood="gaussian", seed=2):
    """Host-side (numpy) parameter point of BASELINE.json configs[2]: basis, PolynomialDecay eigenvalues at
    beta = 1 with the default truncation_level = 10 clip (spherical.py:563-587), q = N(mu, L L^T)."""
    rng = np.random.default_rng(seed)
    V, L, m = synthetic_basis(d, F, T, rng)
    M = sum(m)
    dim = d + 1
    alpha = (dim - 2.0) / 2.0
    tl = 10
    n = np.arange(tl, dtype=np.float64)
    c1 = np.array([_gegenbauer_np(k, alpha, np.array(1.0)) for k in range(tl)])
    eigv = (1 + n) ** (-1.0) * (alpha / (n + alpha) / c1)
    eigv = eigv / np.sum((1 + n) ** (-1.0))
    eig = eigv[np.clip(np.arange(F), 0, tl - 1)]
    mu = 0.1 * rng.standard_normal((M, P))
    Lq = np.stack([np.tril(np.eye(M) + 0.01 * rng.standard_normal((M, M))) for _ in range(P)])
    host = dict(w=np.ones(d), b=np.array(1.0), v=np.array(1.0), eig=eig, V=V, L=L, mu=mu, Lq=Lq, sigma2=np.array(0.1),
                alpha=alpha, order=1, dim=dim, lik=likelihood)
    return host, m


DATA_BLOCK = 65536  # rows per independently seeded block of the global data set


def global_rows(lo, hi, d=8, P=1, likelihood="gaussian", seed=2):
    """Rows [lo, hi) of THE global synthetic data set of a configuration, as CPU float64 tensors (X [n, d], Y [n, P]).

    The data set is defined block by block: rows [k B, (k + 1) B), B = DATA_BLOCK, are drawn from a CPU generator
    seeded with (seed, k) alone, so that row i is the same numbers whatever the world size, the rank that owns it, the
    device or the arm of bench.py that asks (GPU arm, CPU reference arm, tests).  X ~ N(0, I),
    y_j = (j + 1) sum_c sin(x_c) / sqrt(d) + 0.1 eps  (thresholded at 0 for the Bernoulli likelihood)."""
    lo, hi = int(lo), int(hi)
    n = max(hi - lo, 0)
    X = torch.empty((n, d), dtype=torch.float64)
    Y = torch.empty((n, P), dtype=torch.float64)
    if n == 0:
        return X, Y
    g = torch.Generator(device="cpu")
    for k in range(lo // DATA_BLOCK, (hi - 1) // DATA_BLOCK + 1):
        g.manual_seed((int(seed) + 1000) * 1_000_003 + k)
        G = torch.randn((DATA_BLOCK, d + P), dtype=torch.float64, generator=g)
        b0 = k * DATA_BLOCK
        s0, s1 = max(lo, b0) - b0, min(hi, b0 + DATA_BLOCK) - b0
        Xb = G[s0:s1, :d]
        f = torch.sin(Xb).sum(1, keepdim=True) / math.sqrt(d)
        Yb = f * torch.arange(1, P + 1, dtype=torch.float64) + 0.1 * G[s0:s1, d:]
        if likelihood == "bernoulli":
            Yb = (Yb > 0).to(torch.float64)
        X[b0 + s0 - lo:b0 + s1 - lo] = Xb
        Y[b0 + s0 - lo:b0 + s1 - lo] = Yb
    return X, Y


def global_rows_device(lo, hi, device, d=8, P=1, likelihood="gaussian", seed=2, piece=32 * DATA_BLOCK):
    """`global_rows` materialised on `device`, generated and uploaded in pieces of <= 2 M rows (bounded host memory)."""
    lo, hi = int(lo), int(hi)
    X = torch.empty((max(hi - lo, 0), d), dtype=torch.float64, device=device)
    Y = torch.empty((max(hi - lo, 0), P), dtype=torch.float64, device=device)
    for a in range(lo, hi, piece):
        b = min(hi, a + piece)
        Xc, Yc = global_rows(a, b, d, P, likelihood, seed)
        X[a - lo:b - lo].copy_(Xc)
        Y[a - lo:b - lo].copy_(Yc)
    return X, Y


def headline_host_problem(N, d=8, F=11, T=30, P=1, likelihood="gaussian", seed=2):
    """CPU-only variant (numpy rows [0, N) of the global data set) for the reference arm of bench.py: (host params, X, Y, m)."""
    host, m = headline_params(d, F, T, P, likelihood, seed)
    X, Y = global_rows(0, N, d, P, likelihood, seed)
    return host, X.numpy(), Y.numpy(), m


def headline_problem(N, d=8, F=11, T=30, P=1, likelihood="gaussian", seed=2, device="cuda", rank=0, world=1, precision="f64"):
    """BASELINE.json configs[2] (and, through d / F / T / likelihood, configs[3], [4]): full SVGP ELBO+gradient step.
    Rank `rank` of `world` owns the contiguous block `shard_bounds(N, rank, world)` of the N global rows
    (`global_rows`: the same data set for every world size)."""
    from .distributed import shard_bounds
    host, m = headline_params(d, F, T, P, likelihood, seed)
    M, dim = sum(m), d + 1
    lo, hi = shard_bounds(N, rank, world)
    X, Y = global_rows_device(lo, hi, device, d, P, likelihood, seed)
    hp = HotPath(d, m, num_outputs=P, kernel_order=1, likelihood=likelihood, q_kind="tril", weight_mode="ard", device=device,
                 precision=precision)
    vhat = np.concatenate(host["V"], 0)
    lcat = np.concatenate([l.ravel() for l in host["L"]])
    dev = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64), device=device)
    args = (X, Y, dev(host["w"]), dev(host["b"]), dev(host["v"]), dev(host["eig"]), dev(vhat), dev(lcat), dev(host["mu"]),
            dev(host["Lq"]), dev(host["sigma2"]))
    return {
        "hot": hp, "args": args, "m": m, "M": M, "dim": dim, "host": host, "rows": (lo, hi),
        "alg_flops_per_point": algorithmic_flops_per_point(dim, m, True),
        "alg_bytes_per_point": 8 * (d + P),
    }
