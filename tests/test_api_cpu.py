"""Host side of the gpfy API mirror, on the CPU: the structure / mask / invariant tests of the reference's
own suite (tests/test_spherical_harmonics.py, test_variational.py, test_param.py, test_spherical.py) plus
value checks of every N-independent piece against the reference goldens and the oracle."""
import numpy as np
import pytest
import torch

from gpfy_b200 import optimization as opt
from gpfy_b200.param import Exp, FillTriangular, Identity, Param
from gpfy_b200.spherical import NTK, ArcCosine, PolynomialDecay
from gpfy_b200.spherical_harmonics import SphericalHarmonics, _num_harmonics
from gpfy_b200.variational import VariationalDistributionFullCovariance, VariationalDistributionTriL
from oracle import gpfy_oracle as O
from tests.api_util import objects_from_golden
from tests.golden_util import CASES, golden_pack, load, rel_err

API_CASES = list(CASES)


# ---- reference tests/test_spherical_harmonics.py ----
def test_init_raises_on_sphere_mismatch():  # :30-34
    kernel = NTK(depth=3)
    param = kernel.init(0, 5)
    with pytest.raises(ValueError):
        SphericalHarmonics(4, phase_truncation=8).init(0, 3, param)


@pytest.mark.parametrize("input_dim,F,T", [(2, 8, 2**31 - 1), (5, 5, 10), (8, 11, 30), (32, 4, 16)])
def test_param_keys_trainables_and_nan_sentinel(input_dim, F, T):  # :37-62, :83-98
    sh = SphericalHarmonics(F, phase_truncation=T)
    param = sh.init(42, input_dim)
    dim = input_dim + 1
    feats = param.params["variational"]["inducing_features"]
    assert sorted(feats) == [f"V_{n:02d}" for n in range(F)]
    tr = param._trainables["variational"]["inducing_features"]
    has_file = dim <= 20
    for n in range(F):
        nh = _num_harmonics(dim, n)
        assert feats[f"V_{n:02d}"].shape == (min(T, nh), dim)
        assert tr[f"V_{n:02d}"] == (not (nh <= T and has_file))
    ob = param.constants["variational"]["inducing_features"]["orthogonal_basis"]
    assert all(np.all(np.isnan(b)) for b in ob) == any(tr.values())
    assert all(b.shape == (feats[f"V_{n:02d}"].shape[0],) * 2 for n, b in enumerate(ob))
    assert all(v.shape[0] <= T for v in sh.Vs(param))
    assert sh.num_inducing(param) == sum(v.shape[0] for v in feats.values())
    for v in sh.Vs(param):  # :72-80 unit rows
        np.testing.assert_allclose(np.linalg.norm(v, axis=1), 1.0, atol=1e-12)
    assert param.constants["sphere"]["alpha"] == (dim - 2.0) / 2.0
    kernel = PolynomialDecay()
    kp = kernel.init(1, input_dim)
    full = sh.init(3, input_dim, kp)
    assert sh.Kuu(full, kernel).shape == (sh.num_inducing(full),)  # :101-110


def test_missing_untruncated_degree_raises_like_reference():
    with pytest.raises(ValueError):  # fs_9D.npz stops at degree 5 (harmonics/fund_set.py:36)
        SphericalHarmonics(8, phase_truncation=10**6).init(0, 8)
    with pytest.raises(TypeError):
        SphericalHarmonics()


@pytest.mark.parametrize("name", API_CASES)
def test_Vs_Ls_Kuu_eigenvalues_match_reference_golden(name):
    g = load(name)
    o = objects_from_golden(g)
    sh, param, kernel = o["sh"], o["param"], o["kernel"]
    for n, (v, L) in enumerate(zip(sh.Vs(param), sh.Ls(param))):
        assert rel_err(v, g[f"o_Vs_{n:02d}"]) < 1e-14
        assert rel_err(L, g[f"o_Ls_{n:02d}"]) < 1e-10
    assert rel_err(kernel.eigenvalues(param, sh.levels), g["o_eigenvalues"]) < 1e-12
    assert rel_err(sh.Kuu(param, kernel), g["o_Kuu"]) < 1e-12
    assert sh._learned(param) == bool(g["meta_orth_basis_is_nan"])
    assert list(o["init_param"]._trainables["variational"]["inducing_features"].values()) == list(g["meta_trainable_V"])


def test_kernel_init_eigenvalues_match_reference_constants():
    g = load("d5_ntk3_P2")
    param = NTK(depth=3).init(0, 5)
    assert rel_err(param.constants["NTK"]["eigenvalues"], g["const_eigenvalues"]) < 1e-11
    # PolynomialDecay eigenvalues weighted by multiplicity normalise (reference tests/test_spherical.py:145-154)
    k = PolynomialDecay(truncation_level=8)
    p = k.init(0, 4)
    eig = k.eigenvalues(p, np.arange(8))
    np.testing.assert_allclose(sum(eig[n] * _num_harmonics(5, n) for n in range(8)), 1.0, rtol=1e-12)
    with pytest.raises(ValueError):
        ArcCosine(order=3)


# ---- reference tests/test_variational.py ----
@pytest.mark.parametrize("cls", [VariationalDistributionTriL, VariationalDistributionFullCovariance])
def test_variational_init_shapes_logdet_trace(cls):  # :33-75
    q = cls()
    param = q.init(7, 3)
    assert q.mu(param).shape == (7, 3) and q.cov_part(param).shape == (3, 7, 7)
    assert sorted(param.params["variational"][q.name]) == ["Sigma", "mu"]
    np.testing.assert_allclose(q.logdet(param), 0.0, atol=1e-15)
    np.testing.assert_allclose(q.trace(param), 7.0)
    sh = SphericalHarmonics(3, phase_truncation=4)
    with pytest.raises(ValueError):
        q.init(5, 1, sh.init(0, 3))


@pytest.mark.parametrize("name", ["d4_fullcov_fixedV"])
def test_full_covariance_KL_matches_reference_golden(name):
    """The host mirrors of the reference's methods (logdet / trace / _dkl_dcov, variational.py:375-399); the API's
    prior_KL_and_grad runs them on the device (gpfy_prior_kl_valgrad_ws, tests/test_gpu_parity.py)."""
    g = load(name)
    o = objects_from_golden(g)
    q, param = o["q"], o["param"]
    mu, S = q.mu(param), q.cov_part(param)
    kl = -0.5 * float(np.size(mu)) - 0.5 * float(np.sum(q.logdet(param))) + 0.5 * float(np.sum(np.square(mu))) + 0.5 * float(np.sum(q.trace(param)))
    dmu, dsig = mu.copy(), q._dkl_dcov(S)
    assert abs(kl - float(g["o_KL"])) <= 1e-12 * abs(float(g["o_KL"]))
    mu_t, S_t = O.T(g["p_mu"]).requires_grad_(True), O.T(g["p_Sigma"]).requires_grad_(True)
    O.prior_KL(mu_t, S_t, tril=False).backward()
    assert rel_err(dmu, mu_t.grad.numpy()) < 1e-12 and rel_err(dsig, S_t.grad.numpy()) < 1e-10


# ---- reference tests/test_param.py (subset that pins behaviour) ----
def test_param_bijectors_roundtrip_and_fill_triangular_order():
    ft = FillTriangular()
    x = np.arange(1.0, 7.0)
    L = ft.forward(x)
    # TFP FillTriangular: [1..6] -> [[4,0,0],[6,5,0],[3,2,1]]
    np.testing.assert_array_equal(L, np.array([[4.0, 0, 0], [6, 5, 0], [3, 2, 1]]))
    np.testing.assert_array_equal(ft.inverse(L), x)
    np.testing.assert_array_equal(ft.forward_vjp(x, L), x)
    p = Param(params={"k": {"a": 2.0, "b": np.ones(3)}}, _bijectors={"k": {"b": Identity()}})
    assert isinstance(p._bijectors["k"]["a"], Exp) and p._trainables["k"]["a"] is True
    free = p.unconstrained()
    assert not free._constrained and np.isclose(free.params["k"]["a"], np.log(2.0))
    np.testing.assert_allclose(free.constrained().params["k"]["a"], 2.0)
    assert p.params["k"]["a"].dtype == np.float64
    with pytest.raises(ValueError):
        Param(params={"k": {"a": 1.0}}, _trainables={"k": {"zz": False}})
    with pytest.raises(ValueError):
        p.replace_param("nope", a=1.0)
    assert p.set_trainable("k", a=False)._trainables["k"]["a"] is False
    assert float(p.replace_param("k", a=3.0).params["k"]["a"]) == 3.0


# ---- host chain rules against oracle autograd ----
@pytest.mark.parametrize("name", ["d8_F6_T12", "d32_F4_T16_bernoulli"])
def test_inducing_chain_matches_autograd_through_normalise_gram_cholesky(name):
    g = load(name)
    o = objects_from_golden(g)
    sh, param = o["sh"], o["param"]
    F = int(g["meta_F"])
    alpha = float(g["meta_alpha"])
    rng = np.random.default_rng(0)
    Vt = [O.T(g[f"p_V_{n:02d}"]).requires_grad_(True) for n in range(F)]
    Vh = O.normalise_V(Vt)
    Ls = O.orthogonalise_basis(Vh, alpha)
    dvh = [rng.standard_normal(v.shape) for v in Vt]
    dL = [np.tril(rng.standard_normal(l.shape)) for l in Ls]
    s = sum((v * O.T(d)).sum() for v, d in zip(Vh, dvh)) + sum((l * O.T(d)).sum() for l, d in zip(Ls, dL))
    grads = torch.autograd.grad(s, Vt)
    out = opt._inducing_chain(sh, param, np.concatenate(dvh, 0), np.concatenate([d.ravel() for d in dL]))
    for n in range(F):
        assert rel_err(out[f"V_{n:02d}"], grads[n].numpy()) < 1e-9, n


# ---- reference likelihoods.py:172-186, 236-252 ----
def test_conditional_distribution_objects():
    """`Likelihood.conditional_distribution` returns p(Y | F): N(F, sigma^2) for Gaussian, Bernoulli(inv_probit(F)) otherwise -
    host-side stand-ins for the TFP objects (mean / variance / log_prob / sample), checked against the closed forms."""
    from scipy.stats import norm
    from gpfy_b200.likelihoods import Bernoulli, Gaussian
    rng = np.random.default_rng(0)
    F, Y = rng.standard_normal((7, 2)), rng.standard_normal((7, 2))
    lik = Gaussian()
    param = lik.init()
    param.params["likelihood"]["Gaussian"]["variance"] = np.array(0.37)
    d = lik.conditional_distribution(param, F)
    np.testing.assert_allclose(d.mean(), F)
    np.testing.assert_allclose(d.variance(), np.full(F.shape, 0.37))
    np.testing.assert_allclose(d.log_prob(Y), norm.logpdf(Y, F, np.sqrt(0.37)), rtol=1e-13)
    assert d.sample((3,), seed=1).shape == (3, 7, 2)
    b = Bernoulli().conditional_distribution(Bernoulli().init(), F)
    p = 0.5 * (1.0 + torch.erf(torch.as_tensor(F) / 2.0 ** 0.5).numpy()) * (1 - 2e-3) + 1e-3
    np.testing.assert_allclose(b.mean(), p, rtol=1e-13)
    Yb = (Y > 0).astype(float)
    np.testing.assert_allclose(b.log_prob(Yb), Yb * np.log(p) + (1 - Yb) * np.log1p(-p), rtol=1e-13)
    assert set(np.unique(b.sample((5,), seed=2))) <= {0.0, 1.0}
