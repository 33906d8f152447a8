"""Generates tests/golden/*.npz by executing the UNMODIFIED reference source.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

`/root/reference/src/gpfy` is imported as-is; the `jax`, `tensorflow_probability` and `optax`
imports it performs resolve to the numpy stand-ins in `oracle/jax_shim/` (JAX cannot be installed in
this image).  Every case stores the constrained parameters it used (the shim's `jax.random` is not
threefry, so initial V differ from real JAX) together with the reference's outputs for each hot-path
function, plus central finite-difference directional derivatives of the reference `elbo`, which pin
the oracle's autograd gradients.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "jax_shim"))
sys.path.insert(0, "/root/reference/src")

import jax  # noqa: E402  (the shim)
import jax.numpy as jnp  # noqa: E402
from gpfy.gegenbauer import GegenbauerLookupTable, gegenbauer  # noqa: E402
from gpfy.harmonics.utils import funk_hecke_lambda  # noqa: E402
from gpfy.likelihoods import Bernoulli, Gaussian  # noqa: E402
from gpfy.model import GP  # noqa: E402
from gpfy.optimization import elbo  # noqa: E402
from gpfy.spherical import NTK, ArcCosine, PolynomialDecay  # noqa: E402
from gpfy.spherical_harmonics import SphericalHarmonics, _num_harmonics  # noqa: E402
from gpfy.variational import VariationalDistributionFullCovariance, VariationalDistributionTriL  # noqa: E402

assert jax.__file__.startswith(os.path.join(ROOT, "oracle", "jax_shim")), jax.__file__


def copy_tree(t):
    if isinstance(t, dict):
        return {k: copy_tree(v) for k, v in t.items()}
    return np.array(t, dtype=np.float64)


def build_case(name, kernel, F, Tr, input_dim, N, lik, q, P=1, projection_dim=None, seed=0, full_cov=False):
    rng = np.random.default_rng(seed)
    key = jax.random.PRNGKey(seed)
    sh = SphericalHarmonics(F, phase_truncation=Tr)
    m = GP(kernel).conditional(sh, q)
    param = m.init(key, input_dim=input_dim, num_independent_processes=P, projection_dim=projection_dim,
                   likelihood=lik, sh_features=sh, variational_dist=q)
    M = int(sh.num_inducing(param))
    p = copy_tree(param.params)
    kp = p[kernel.name]
    if projection_dim:
        kp["weight_variances"] = 0.5 * rng.standard_normal((input_dim, projection_dim))
    elif not kernel.ard:
        kp["weight_variances"] = np.array(0.8)     # one shared weight variance (spherical.py:110-114)
    else:
        kp["weight_variances"] = np.exp(0.3 * rng.standard_normal(input_dim))
    kp["bias_variance"] = np.array(0.7)
    kp["variance"] = np.array(1.3)
    if "beta" in kp:
        kp["beta"] = np.array(1.4)
    p["variational"][q.name]["mu"] = 0.3 * rng.standard_normal((M, P))
    if full_cov:
        S = []
        for _ in range(P):
            Lr = np.tril(np.eye(M) + 0.05 * rng.standard_normal((M, M)))
            S.append(Lr @ Lr.T)
        p["variational"][q.name]["Sigma"] = np.stack(S)
    else:
        p["variational"][q.name]["Sigma"] = np.stack(
            [np.tril(np.eye(M) + 0.05 * rng.standard_normal((M, M))) for _ in range(P)])
    if "likelihood" in p and "Gaussian" in p["likelihood"]:
        p["likelihood"]["Gaussian"]["variance"] = np.array(0.3)
    param = param.replace(params=p)

    X = rng.standard_normal((N, input_dim))
    f = np.stack([np.sin(X[:, : min(3, input_dim)] * (j + 1)).sum(-1) for j in range(P)], -1)
    if isinstance(lik, Bernoulli):
        Y = (f + 0.3 * rng.standard_normal(f.shape) > 0).astype(np.float64)
    else:
        Y = f + 0.1 * rng.standard_normal(f.shape)

    out = {}
    out["meta_F"] = np.array(F)
    out["meta_T"] = np.array(min(Tr, 2**31 - 1))
    out["meta_input_dim"] = np.array(input_dim)
    out["meta_P"] = np.array(P)
    out["meta_order"] = np.array(kernel.order)
    out["meta_projection_dim"] = np.array(projection_dim or 0)
    out["meta_kernel"] = np.array(kernel.name)
    out["meta_ard"] = np.array(bool(kernel.ard))
    out["meta_lik"] = np.array(lik.name)
    out["meta_q"] = np.array(q.name)
    out["meta_alpha"] = np.array(param.constants["sphere"]["alpha"])
    if isinstance(kernel, PolynomialDecay):
        out["meta_truncation_level"] = np.array(kernel.truncation_level)
    else:
        out["meta_depth"] = np.array(kernel.depth)
        out["const_eigenvalues"] = np.array(param.constants[kernel.name]["eigenvalues"])
    out["X"], out["Y"] = X, Y
    for k, v in kp.items():
        out[f"p_{k}"] = np.array(v)
    out["p_mu"] = p["variational"][q.name]["mu"]
    out["p_Sigma"] = p["variational"][q.name]["Sigma"]
    if "likelihood" in p and "Gaussian" in p["likelihood"]:
        out["p_sigma2"] = p["likelihood"]["Gaussian"]["variance"]
    Vraw = [p["variational"]["inducing_features"][f"V_{n:02d}"] for n in range(F)]
    tr = [param._trainables["variational"]["inducing_features"][f"V_{n:02d}"] for n in range(F)]
    ob = param.constants["variational"]["inducing_features"]["orthogonal_basis"]
    out["meta_trainable_V"] = np.array(tr)
    out["meta_orth_basis_is_nan"] = np.array(bool(np.isnan(np.sum(ob[0]))))
    for n in range(F):
        out[f"p_V_{n:02d}"] = Vraw[n]
        out[f"c_orth_{n:02d}"] = np.array(ob[n])

    # ---- reference outputs of every hot-path function ----
    Vs = sh.Vs(param)
    Ls = sh.Ls(param)
    for n in range(F):
        out[f"o_Vs_{n:02d}"] = np.array(Vs[n])
        out[f"o_Ls_{n:02d}"] = np.array(Ls[n])
    out["o_eigenvalues"] = np.array(kernel.eigenvalues(param, sh.levels))
    xs, r = kernel.to_sphere(param, X)
    out["o_x_sphere"], out["o_r"] = np.array(xs), np.array(r)
    out["o_K_diag"] = np.array(kernel.K_diag(param, X))
    out["o_Phi"] = np.array(sh.polynomial_expansion(param, xs))
    out["o_Kuf"] = np.array(sh.Kuf(param, kernel, X))
    out["o_Kuu"] = np.array(sh.Kuu(param, kernel))
    proj_fun = m.conditional_fn[0]
    A = proj_fun(param, X)
    out["o_proj"] = np.array(A)
    out["o_project_mean"] = np.array(q.project_mean(param, A))
    out["o_project_diag_variance"] = np.array(q.project_diag_variance(param, A))
    fmu, fvar = m.predict_diag(param, X)
    out["o_fmu"], out["o_fvar"] = np.array(fmu), np.array(fvar)
    out["o_var_exp"] = np.array(lik.variational_expectations(param, fmu, fvar)(Y))
    out["o_KL"] = np.array(q.prior_KL(param))
    out["o_elbo"] = np.array(elbo(param, m, q, lik, (X, Y)))
    out["o_elbo_ds"] = np.array(elbo(param, m, q, lik, (X, Y), dataset_size=10 * N))

    # ---- the kernel-matrix side: K, K2, shape_function, full predictive covariance ----
    X2 = X[:7] * 0.9 + 0.1
    out["X2"] = X2
    out["o_K"] = np.array(kernel.K(param, X))
    out["o_K_cross"] = np.array(kernel.K(param, X, X2))
    out["o_K2_cross"] = np.array(kernel.K2(param, X, X2))
    out["shape_x"] = np.linspace(-1.0, 1.0, 41)
    out["o_shape"] = np.array(kernel.shape_function(param, jnp.asarray(out["shape_x"])))
    out["o_project_variance"] = np.array(q.project_variance(param, A))
    pm, pc = m.predict(param, X)
    out["o_predict_cov"] = np.array(pc)

    # ---- central finite differences of the reference elbo along stored random directions ----
    def with_leaf(path, val):
        pp = copy_tree(param.params)
        d = pp
        for k in path[:-1]:
            d = d[k]
        d[path[-1]] = val
        return param.replace(params=pp)

    leaves = {
        "mu": ("variational", q.name, "mu"),
        "Sigma": ("variational", q.name, "Sigma"),
        "w": (kernel.name, "weight_variances"),
        "b": (kernel.name, "bias_variance"),
        "v": (kernel.name, "variance"),
    }
    if "beta" in kp:
        leaves["beta"] = (kernel.name, "beta")
    if "p_sigma2" in out:
        leaves["sigma2"] = ("likelihood", "Gaussian", "variance")
    for n in range(F):
        if tr[n]:
            leaves[f"V_{n:02d}"] = ("variational", "inducing_features", f"V_{n:02d}")
    h = 1e-6
    for lname, path in leaves.items():
        base = param.params
        for k in path:
            base = base[k]
        base = np.array(base, dtype=np.float64)
        dirn = rng.standard_normal(base.shape)
        if lname == "Sigma":
            dirn = np.tril(dirn) if not full_cov else dirn + np.swapaxes(dirn, -1, -2)
        fp = float(elbo(with_leaf(path, base + h * dirn), m, q, lik, (X, Y)))
        fm = float(elbo(with_leaf(path, base - h * dirn), m, q, lik, (X, Y)))
        out[f"fd_dir_{lname}"] = dirn
        out[f"fd_val_{lname}"] = np.array((fp - fm) / (2 * h))

    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)
    print(f"{name}: M={M} N={N} elbo={float(out['o_elbo']):.6f}")


def build_gegenbauer():
    """Gegenbauer recurrence / derivative / lookup / Funk-Hecke outputs of the reference."""
    rng = np.random.default_rng(7)
    x = np.concatenate([rng.uniform(-1, 1, 61), [-1.0, 0.0, 1.0]])
    out = {"x": x}
    for alpha in (0.5, 1.5, 3.0, 3.5, 10.0, 15.5):
        for n in (0, 1, 2, 5, 10, 12):
            out[f"C_{n}_{alpha}"] = np.array(gegenbauer(n, alpha, x))
            from gpfy.gegenbauer import gegenbauer_fwd
            out[f"dC_{n}_{alpha}"] = np.array(gegenbauer_fwd(n, alpha, x)[1])
    lut = GegenbauerLookupTable(10, 3.5)
    out["lut_3.5_C5"] = np.array(lut(5, 3.5, x))
    out["lut_3.5_C10_at1"] = np.array(lut(10, 3.5, jnp.array(1.0)))
    ntk = NTK(depth=3)
    from gpfy.param import Param
    out["fh_ntk3_d6"] = np.array([float(funk_hecke_lambda(lambda t: ntk.shape_function(Param(), x=t), n, 6)) for n in range(8)])
    ac = ArcCosine(depth=2, order=1)
    out["fh_arccos2_d9"] = np.array([float(funk_hecke_lambda(lambda t: ac.shape_function(Param(), x=t), n, 9)) for n in range(8)])
    out["num_harmonics"] = np.array([[_num_harmonics(d, n) for n in range(13)] for d in (3, 8, 9, 33)])
    np.savez_compressed(os.path.join(HERE, "gegenbauer.npz"), **out)
    print("gegenbauer: done")


if __name__ == "__main__":
    tril = VariationalDistributionTriL()
    full = VariationalDistributionFullCovariance()
    if "--orders" in sys.argv:   # only the kernel-order cases (adds to an existing set of goldens)
        build_case("d3_arccos_order0", ArcCosine(depth=1, order=0), 4, 6, 3, 24, Gaussian(), tril, seed=9)
        build_case("d3_arccos_order2_bern", ArcCosine(depth=1, order=2), 4, 6, 3, 24, Bernoulli(), tril, P=2, seed=10)
        build_case("d4_ntk3_scalar_w", NTK(depth=3, ard=False), 4, 8, 4, 24, Gaussian(), tril, seed=11)
        sys.exit(0)
    build_gegenbauer()
    # config 1 of BASELINE.json: S^2 (input_dim 2 => dim 3), degrees 0..10, no truncation
    build_case("s2_deg10_polydecay", PolynomialDecay(truncation_level=11), 11, 2**31 - 1, 2, 48, Gaussian(), tril, seed=1)
    # same with the default truncation_level=10 so that degree 10 is clipped to lambda_9
    build_case("s2_deg10_clip", PolynomialDecay(), 11, 2**31 - 1, 2, 32, Gaussian(), tril, seed=2)
    # configs 2/3 shape at reduced T / N: d=8 => dim 9, phase truncation bites from degree 2
    build_case("d8_F6_T12", PolynomialDecay(), 6, 12, 8, 64, Gaussian(), tril, seed=3)
    # headline basis: d=8, degrees 0..10, T=30  (M=280)
    build_case("d8_F11_T30", PolynomialDecay(), 11, 30, 8, 40, Gaussian(), tril, seed=4)
    # the reference's own test fixture kernel: NTK(depth=3) at input_dim=5; two outputs
    build_case("d5_ntk3_P2", NTK(depth=3), 5, 10, 5, 40, Gaussian(), tril, P=2, seed=5)
    # config 5 shape: d=32 => dim 33 (no fundamental system shipped => every V random), Bernoulli
    build_case("d32_F4_T16_bernoulli", PolynomialDecay(), 4, 16, 32, 48, Bernoulli(), tril, seed=6)
    # full-covariance q, fixed V with precomputed orthogonal basis (no NaN sentinel)
    build_case("d4_fullcov_fixedV", PolynomialDecay(), 4, 2**31 - 1, 4, 40, Gaussian(), full, seed=7, full_cov=True)
    # linear projection weights [d, p]
    build_case("d6_proj3_arccos", ArcCosine(depth=2, order=1), 5, 8, 6, 40, Gaussian(), tril, projection_dim=3, seed=8)
    # kernel order 0 and 2 (K_diag = v r^(2 order), spherical.py:400; kappa_0 / kappa_2, :299-309)
    build_case("d3_arccos_order0", ArcCosine(depth=1, order=0), 4, 6, 3, 24, Gaussian(), tril, seed=9)
    build_case("d3_arccos_order2_bern", ArcCosine(depth=1, order=2), 4, 6, 3, 24, Bernoulli(), tril, P=2, seed=10)
    # ard=False: a single weight variance shared by all input dimensions
    build_case("d4_ntk3_scalar_w", NTK(depth=3, ard=False), 4, 8, 4, 24, Gaussian(), tril, seed=11)
