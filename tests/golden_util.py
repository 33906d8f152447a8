"""Loads tests/golden/*.npz (reference outputs, see tests/golden/make_golden.py) into hot-path packs."""
import glob
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))
               if not p.endswith("gegenbauer.npz"))


def load(name):
    with np.load(os.path.join(GOLDEN_DIR, name + ".npz")) as f:
        return {k: f[k] for k in f.files}


def golden_pack(g, eig=None):
    """Constrained hot-path pack (see oracle/gpfy_oracle.py) from a golden case.

    `eig` overrides the per-degree eigenvalues (default: the reference's `kernel.eigenvalues` output)."""
    F = int(g["meta_F"])
    learned = bool(g["meta_orth_basis_is_nan"])
    pack = {
        "dim": int(g["p_V_00"].shape[1]),
        "alpha": float(g["meta_alpha"]),
        "V": [g[f"p_V_{n:02d}"] for n in range(F)],
        "L": None if learned else [g[f"c_orth_{n:02d}"] for n in range(F)],
        "w": g["p_weight_variances"],
        "b": g["p_bias_variance"],
        "v": g["p_variance"],
        "eig": g["o_eigenvalues"] if eig is None else eig,
        "order": int(g["meta_order"]),
        "mu": g["p_mu"],
        "lik": "gaussian" if str(g["meta_lik"]) == "Gaussian" else "bernoulli",
    }
    if str(g["meta_q"]).endswith("TriL"):
        pack["Lq"] = g["p_Sigma"]
    else:
        pack["Sigma"] = g["p_Sigma"]
    if "p_sigma2" in g:
        pack["sigma2"] = g["p_sigma2"]
    # shape-function description (kernel-matrix side only)
    pack["kernel"] = str(g["meta_kernel"])
    if pack["kernel"] == "PolynomialDecay":
        pack["truncation_level"] = int(g["meta_truncation_level"])
        pack["beta"] = g["p_beta"]
    else:
        pack["depth"] = int(g["meta_depth"])
    return pack


def rel_err(a, b):
    """max |a-b| / max(|b|_inf, tiny): the tolerance convention used by every parity test."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    if a.shape != b.shape:
        raise AssertionError(f"shape mismatch {a.shape} vs {b.shape}")
    if a.size == 0:
        return 0.0
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def elem_err(a, b, floor=1e-6):
    """Componentwise: max_i |a_i - b_i| / max(|b_i|, floor * |b|_inf) - the per-element relative error of every entry that is
    not smaller than `floor` times the largest one (below that the error is measured against the floor)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    if a.shape != b.shape:
        raise AssertionError(f"shape mismatch {a.shape} vs {b.shape}")
    if a.size == 0:
        return 0.0
    den = np.maximum(np.abs(b), max(floor * np.max(np.abs(b)), 1e-300))
    return float(np.max(np.abs(a - b) / den))
