"""Known-answer / invariant tests the reference's own tests hold for the hot path, run on the oracle.

Each test cites the reference test it restates (paths relative to /root/reference)."""
import math

import numpy as np
import pytest
import scipy.special
import torch
from scipy.integrate import quad

from oracle import gpfy_oracle as O
from tests.golden_util import golden_pack, load

ALPHAS = [0.5, 1.5, 3.0, 10.0]
FREQS = [0, 1, 2, 5, 10]


@pytest.mark.parametrize("n", FREQS)
@pytest.mark.parametrize("alpha", ALPHAS)
@pytest.mark.parametrize("shape", [(), (1, 1), (3, 1), (100, 7)])
def test_gegenbauer_vs_scipy(n, alpha, shape):
    """tests/test_gegenbauer.py:36-43 (made to actually assert)."""
    x = np.random.default_rng(0).uniform(-1, 1, shape)
    ours = O._gegenbauer(n, alpha, O.T(x)).numpy()
    ref = scipy.special.gegenbauer(n, alpha)(x)
    np.testing.assert_allclose(ours, ref, rtol=1e-10, atol=1e-10 * max(1.0, float(np.max(np.abs(ref)))))


@pytest.mark.parametrize("n", FREQS)
@pytest.mark.parametrize("alpha", ALPHAS)
def test_gegenbauer_custom_gradient_is_the_derivative(n, alpha):
    """tests/test_gegenbauer.py:56-64: custom VJP == d/dx (here vs central differences)."""
    x = np.linspace(-0.95, 0.95, 41)
    h = 1e-6
    fd = (O._gegenbauer(n, alpha, O.T(x + h)) - O._gegenbauer(n, alpha, O.T(x - h))).numpy() / (2 * h)
    an = O.gegenbauer_grad(n, alpha, O.T(x)).numpy()
    np.testing.assert_allclose(an, fd, rtol=1e-6, atol=1e-6 * max(1.0, float(np.max(np.abs(fd)))))


def test_gegenbauer_at_one_closed_form():
    """C_n^a(1) = Gamma(n + 2a) / (Gamma(2a) n!)  (SURVEY §8c iii); 3003 at a=3, n=10."""
    assert abs(float(O._gegenbauer(10, 3.0, O.T(1.0))) - 3003.0) < 1e-9
    for alpha in (0.5, 3.0, 3.5, 15.5):
        for n in range(0, 13):
            ref = math.exp(math.lgamma(n + 2 * alpha) - math.lgamma(2 * alpha) - math.lgamma(n + 1))
            assert abs(float(O._gegenbauer(n, alpha, O.T(1.0))) / ref - 1) < 1e-12


@pytest.mark.parametrize("n", FREQS)
@pytest.mark.parametrize("alpha", ALPHAS)
def test_gegenbauer_orthogonality(n, alpha):
    """tests/test_utils.py:26-37."""
    dim = int(2 * alpha + 2)
    f = lambda x: float(O._gegenbauer(n, alpha, O.T(x)) * O._gegenbauer(n + 1, alpha, O.T(x))) * (1 - x * x) ** ((dim - 3) / 2)
    res, tol = quad(f, -1, 1)
    assert abs(res) <= max(tol, 1e-8)


def test_num_harmonics_known_values():
    """spherical_harmonics.py:31-47; SURVEY §8 size table."""
    assert [O.num_harmonics(3, n) for n in range(11)] == [2 * n + 1 for n in range(11)]
    assert [O.num_harmonics(9, n) for n in range(11)] == [1, 9, 44, 156, 450, 1122, 2508, 5148, 9867, 17875, 30888]
    assert [O.num_harmonics(33, n) for n in range(5)] == [1, 33, 560, 6512, 58344]


@pytest.mark.parametrize("name", ["d8_F6_T12", "d5_ntk3_P2", "d32_F4_T16_bernoulli"])
def test_learned_Vs_are_unit_norm(name):
    """tests/test_spherical_harmonics.py:72-80."""
    pack = O.pack_to_torch(golden_pack(load(name)))
    Vhat, _ = O.resolve_V_L(pack)
    for v in Vhat:
        np.testing.assert_allclose(torch.sum(v * v, 1).numpy(), 1.0, rtol=1e-14)


def test_truncation_mask_and_nan_sentinel():
    """tests/test_spherical_harmonics.py:37-62,83-89: trainable iff N(dim, n) > T; NaN basis iff any trainable."""
    g = load("d8_F6_T12")
    T_, F = int(g["meta_T"]), int(g["meta_F"])
    dim = g["p_V_00"].shape[1]
    for n in range(F):
        assert bool(g["meta_trainable_V"][n]) == (O.num_harmonics(dim, n) > T_)
        assert g[f"p_V_{n:02d}"].shape[0] == min(T_, O.num_harmonics(dim, n))
    assert bool(g["meta_orth_basis_is_nan"]) == bool(g["meta_trainable_V"].any())
    g2 = load("d4_fullcov_fixedV")
    assert not g2["meta_trainable_V"].any() and not bool(g2["meta_orth_basis_is_nan"])


@pytest.mark.parametrize("name", ["s2_deg10_polydecay", "d4_fullcov_fixedV"])
def test_addition_theorem_complete_degrees(name):
    """for a complete degree, sum_m Phi_lm(x)^2 = N(dim, l) for every unit x."""
    g = load(name)
    pack = O.pack_to_torch(golden_pack(g))
    Vhat, L = O.resolve_V_L(pack)
    xs, _ = O.to_sphere(pack["w"], pack["b"], O.T(g["X"]))
    Phi = O.polynomial_expansion(Vhat, L, pack["alpha"], xs).numpy()
    off = 0
    for n, v in enumerate(Vhat):
        m = v.shape[0]
        assert m == O.num_harmonics(pack["dim"], n)
        np.testing.assert_allclose(np.sum(Phi[off:off + m] ** 2, 0), m, rtol=1e-9)
        off += m


def test_to_sphere_invariants():
    """tests/test_spherical.py:103-120: unit norm; |scaled|^2 = r^2 - bias_variance."""
    rng = np.random.default_rng(3)
    X = O.T(rng.standard_normal((17, 5)))
    w, b = O.T(np.exp(rng.standard_normal(5))), O.T(1.0)
    xs, r = O.to_sphere(w, b, X)
    np.testing.assert_allclose(torch.sum(xs ** 2, -1).numpy(), 1.0, rtol=1e-14)
    np.testing.assert_allclose(torch.sum(O.scale_X(w, X) ** 2, -1).numpy(), (r ** 2 - 1).numpy(), rtol=1e-12)


def test_polydecay_eigenvalues_normalise():
    """tests/test_spherical.py:145-154."""
    alpha = (6 - 2.0) / 2.0
    eig = O.polydecay_eigenvalues(O.T(1.0), alpha, 10, range(10))
    cf = O.polydecay_const_factor(alpha, 10)
    assert abs(float(torch.sum(eig / cf)) - 1.0) < 1e-14


@pytest.mark.parametrize("lik", ["gaussian", "bernoulli"])
def test_var_exp_equals_log_prob_at_zero_variance(lik):
    """tests/test_likelihoods.py:32-46."""
    rng = np.random.default_rng(5)
    f = O.T(rng.standard_normal((10, 5)))
    if lik == "gaussian":
        y = O.T(rng.standard_normal((10, 5)))
        s2 = O.T(1.0)
        r1 = torch.sum(-0.5 * ((y - f) / torch.sqrt(s2)) ** 2 - torch.log(torch.sqrt(s2)) - 0.5 * math.log(2 * math.pi), -1)
        r2 = O.gaussian_var_exp(s2, f, torch.zeros_like(f), y)
    else:
        y = O.T((rng.standard_normal((10, 5)) > 0).astype(float))
        r1 = O.bernoulli_log_prob(f, y)
        r2 = O.bernoulli_var_exp(f, torch.zeros_like(f), y)
    np.testing.assert_allclose(r1.numpy(), r2.numpy(), rtol=1e-12)


def test_q_init_logdet_trace_and_kl_zero():
    """tests/test_variational.py:57-98: at mu=0, L=I: logdet = 0, trace = M, hence KL = 0."""
    M, P = 13, 2
    mu = torch.zeros(M, P, dtype=O.DT)
    L = torch.eye(M, dtype=O.DT).expand(P, M, M).clone()
    assert float(O.prior_KL(mu, L, tril=True)) == 0.0
    assert float(O.prior_KL(mu, L, tril=False)) == 0.0


def test_elbo_at_init_is_independent_of_features():
    """mu = 0, L_q = I => fvar = K_diag exactly and
    ELBO = sum_n [-1/2 log(2 pi s2) - (y^2 + v r^2) / (2 s2)]."""
    g = load("d8_F6_T12")
    pack = golden_pack(g)
    M = pack["mu"].shape[0]
    pack["mu"] = np.zeros((M, 1))
    pack["Lq"] = np.eye(M)[None]
    tp = O.pack_to_torch(pack)
    X, Y = O.T(g["X"]), O.T(g["Y"])
    _, r = O.to_sphere(tp["w"], tp["b"], X)
    s2 = tp["sigma2"]
    ref = torch.sum(-0.5 * torch.log(2 * math.pi * s2) - (Y[:, 0] ** 2 + tp["v"] * r ** 2) / (2 * s2))
    assert abs(float(O.elbo(tp, X, Y)) / float(ref) - 1) < 1e-12
    fmu, fvar = O.predict_diag(tp, X)
    np.testing.assert_allclose(fvar[:, 0].numpy(), O.k_diag(tp["v"], r, 1).numpy(), rtol=1e-10)
    assert float(torch.max(torch.abs(fmu))) == 0.0


def test_autograd_matches_central_differences_of_oracle():
    """SURVEY §8c (iv)."""
    g = load("d8_F6_T12")
    pack = golden_pack(g)
    val, grads = O.elbo_and_grads(pack, g["X"], g["Y"])
    rng = np.random.default_rng(0)
    h = 1e-6
    for key in ("mu", "w", "v", "sigma2"):
        d = rng.standard_normal(np.shape(pack[key]))
        pp, pm = dict(pack), dict(pack)
        pp[key] = pack[key] + h * d
        pm[key] = pack[key] - h * d
        fd = (float(O.elbo(O.pack_to_torch(pp), O.T(g["X"]), O.T(g["Y"]))) - float(O.elbo(O.pack_to_torch(pm), O.T(g["X"]), O.T(g["Y"])))) / (2 * h)
        an = float(np.sum(grads[key] * d))
        assert abs(an - fd) <= 2e-6 * max(1.0, abs(fd))
