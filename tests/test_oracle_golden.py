"""Pins oracle/gpfy_oracle.py to outputs of the unmodified reference source (tests/golden/*.npz).

Tolerance: 1e-12 relative to the infinity norm of each reference array (both sides are IEEE float64
evaluations of the same formulas; the slack covers BLAS summation order).  Finite-difference
directional derivatives of the reference ELBO pin the oracle's autograd gradients at 2e-6.
"""
import numpy as np
import pytest
import torch

from oracle import gpfy_oracle as O
from tests.golden_util import CASES, golden_pack, load, rel_err

TOL = 1e-12


@pytest.mark.parametrize("name", CASES)
def test_forward_functions_match_reference(name):
    g = load(name)
    pack = O.pack_to_torch(golden_pack(g))
    X, Y = O.T(g["X"]), O.T(g["Y"])
    F = int(g["meta_F"])
    Vhat, L = O.resolve_V_L(pack)
    for n in range(F):
        assert rel_err(Vhat[n].numpy(), g[f"o_Vs_{n:02d}"]) < TOL
        assert rel_err(L[n].numpy(), g[f"o_Ls_{n:02d}"]) < 1e-11
    xs, r = O.to_sphere(pack["w"], pack["b"], X)
    assert rel_err(xs.numpy(), g["o_x_sphere"]) < TOL
    assert rel_err(r.numpy(), g["o_r"]) < TOL
    assert rel_err(O.k_diag(pack["v"], r, pack["order"]).numpy(), g["o_K_diag"]) < TOL
    Phi = O.polynomial_expansion(Vhat, L, pack["alpha"], xs)
    assert Phi.shape == g["o_Phi"].shape
    assert rel_err(Phi.numpy(), g["o_Phi"]) < 1e-11
    assert rel_err(O.Kuf(pack, X).numpy(), g["o_Kuf"]) < 1e-11
    m = [v.shape[0] for v in pack["V"]]
    assert rel_err(O.Kuu(pack["eig"], m).numpy(), g["o_Kuu"]) < TOL
    A = O.projection(pack, X)
    assert rel_err(A.numpy(), g["o_proj"]) < 1e-11
    assert rel_err(O.project_mean(pack["mu"], A).numpy(), g["o_project_mean"]) < 1e-11
    if "Lq" in pack:
        qv = O.project_diag_variance_tril(pack["Lq"], A)
    else:
        qv = O.project_diag_variance_full(pack["Sigma"], A)
    assert rel_err(qv.numpy(), g["o_project_diag_variance"]) < 1e-11
    fmu, fvar = O.predict_diag(pack, X)
    assert rel_err(fmu.numpy(), g["o_fmu"]) < 1e-11
    assert rel_err(fvar.numpy(), g["o_fvar"]) < 1e-11
    assert rel_err(O.var_exp(pack, fmu, fvar, Y).numpy(), g["o_var_exp"]) < 1e-11
    KL = O.prior_KL(pack["mu"], pack.get("Lq", pack.get("Sigma")), tril="Lq" in pack)
    assert rel_err(KL.numpy(), g["o_KL"]) < 1e-11
    assert rel_err(O.elbo(pack, X, Y).numpy(), g["o_elbo"]) < 1e-11
    assert rel_err(O.elbo(pack, X, Y, dataset_size=10 * X.shape[0]).numpy(), g["o_elbo_ds"]) < 1e-11


@pytest.mark.parametrize("name", CASES)
def test_kernel_matrix_side_matches_reference(name):
    """K, K2, shape_function, project_variance and the full predictive covariance."""
    g = load(name)
    pack = O.pack_to_torch(golden_pack(g))
    if "beta" in pack:
        pack["beta"] = O.T(pack["beta"])
    X, X2 = O.T(g["X"]), O.T(g["X2"])
    assert rel_err(O.shape_function(pack, O.T(g["shape_x"])).numpy(), g["o_shape"]) < TOL
    assert rel_err(O.K(pack, X).numpy(), g["o_K"]) < TOL
    assert rel_err(O.K(pack, X, X2).numpy(), g["o_K_cross"]) < TOL
    assert rel_err(O.K(pack, X, X2, unit_diagonal=False).numpy(), g["o_K2_cross"]) < TOL
    A = O.projection(pack, X)
    pv = O.project_variance_tril(pack["Lq"], A) if "Lq" in pack else O.project_variance_full(pack["Sigma"], A)
    assert rel_err(pv.numpy(), g["o_project_variance"]) < 1e-11
    fmu, cov = O.predict(pack, X)
    assert rel_err(fmu.numpy(), g["o_fmu"]) < 1e-11
    assert rel_err(cov.numpy(), g["o_predict_cov"]) < 1e-11


@pytest.mark.parametrize("name", CASES)
def test_eigenvalues_match_reference(name):
    g = load(name)
    F = int(g["meta_F"])
    alpha = float(g["meta_alpha"])
    if str(g["meta_kernel"]) == "PolynomialDecay":
        eig = O.polydecay_eigenvalues(O.T(g["p_beta"]), alpha, int(g["meta_truncation_level"]), range(F))
    else:
        sphere_dim = int(2 * alpha + 1)
        depth = int(g["meta_depth"])
        if str(g["meta_kernel"]) == "NTK":
            fn = lambda t: O.ntk_shape_function(t, depth)
        else:
            fn = lambda t: O.arccosine_shape_function(t, depth, int(g["meta_order"]))
        all_eig = O.spherical_eigenvalues(fn, sphere_dim)
        assert rel_err(all_eig.numpy(), g["const_eigenvalues"]) < 1e-11
        eig = all_eig[:F]
    assert rel_err(eig.numpy(), g["o_eigenvalues"]) < 1e-12


@pytest.mark.parametrize("name", CASES)
def test_autograd_gradients_match_reference_finite_differences(name):
    """(elbo(theta + h d) - elbo(theta - h d)) / 2h of the REFERENCE vs <grad_oracle, d>."""
    g = load(name)
    pack = golden_pack(g)
    F = int(g["meta_F"])
    has_beta = "p_beta" in g
    tp = O.pack_to_torch(pack, requires_grad=True)
    if has_beta:  # eigenvalues are a function of beta (spherical.py:563-587)
        beta = O.T(g["p_beta"]).clone().requires_grad_(True)
        tp["eig"] = O.polydecay_eigenvalues(beta, tp["alpha"], int(g["meta_truncation_level"]), range(F))
    val = O.elbo(tp, O.T(g["X"]), O.T(g["Y"]))
    leaves = {"mu": tp["mu"], "Sigma": tp.get("Lq", tp.get("Sigma")), "w": tp["w"], "b": tp["b"], "v": tp["v"]}
    if has_beta:
        leaves["beta"] = beta
    if "sigma2" in tp:
        leaves["sigma2"] = tp["sigma2"]
    for n in range(F):
        if f"fd_val_V_{n:02d}" in g:
            leaves[f"V_{n:02d}"] = tp["V"][n]
    grads = torch.autograd.grad(val, list(leaves.values()), allow_unused=True)
    checked = 0
    for (lname, _), gr in zip(leaves.items(), grads):
        fd = float(g[f"fd_val_{lname}"])
        an = float(np.sum(gr.numpy() * g[f"fd_dir_{lname}"])) if gr is not None else 0.0
        assert abs(an - fd) <= 2e-6 * max(abs(fd), 1.0), (lname, an, fd)
        checked += 1
    assert checked >= 5


def test_gegenbauer_matches_reference():
    g = load("gegenbauer")
    x = O.T(g["x"])
    for alpha in (0.5, 1.5, 3.0, 3.5, 10.0, 15.5):
        for n in (0, 1, 2, 5, 10, 12):
            assert rel_err(O._gegenbauer(n, alpha, x).numpy(), g[f"C_{n}_{alpha}"]) < 1e-14
            assert rel_err(O.gegenbauer_grad(n, alpha, x).numpy(), g[f"dC_{n}_{alpha}"]) < 1e-14 or n == 0
    # C_10^{3.5}(1) read through the reference's lookup table equals the recurrence at exactly 1.0
    assert rel_err(O._gegenbauer(10, 3.5, O.T(1.0)).numpy(), g["lut_3.5_C10_at1"]) < 1e-15
    nh = np.array([[O.num_harmonics(d, n) for n in range(13)] for d in (3, 8, 9, 33)])
    assert (nh == g["num_harmonics"]).all()
    fn = lambda t: O.ntk_shape_function(t, 3)
    fh = np.array([float(O.funk_hecke_lambda(fn, n, 6)) for n in range(8)])
    assert rel_err(fh, g["fh_ntk3_d6"]) < 1e-11
    fn = lambda t: O.arccosine_shape_function(t, 2, 1)
    fh = np.array([float(O.funk_hecke_lambda(fn, n, 9)) for n in range(8)])
    assert rel_err(fh, g["fh_arccos2_d9"]) < 1e-11
