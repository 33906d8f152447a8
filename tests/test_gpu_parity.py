"""GPU parity: the CUDA path (through the C-ABI) against the CPU oracle on the golden cases.

Tolerance (BASELINE.json north_star): relative 1e-10 in fp64, measured as max|a-b| / max|b| per array
(`rel_err`).  Values the reference produces as exact zeros are compared with an absolute 1e-10.
"""
import numpy as np
import pytest
import torch

from oracle import gpfy_oracle as O
from tests.golden_util import CASES, golden_pack, load, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-10

ARD_CASES = list(CASES)  # ARD, scalar and [d, p] projection weights (spherical.py:110-117, 225-228)


def _resolved(name):
    g = load(name)
    pack = golden_pack(g)
    tp = O.pack_to_torch(pack)
    Vh, L = O.resolve_V_L(tp)
    pack["V"] = [x.numpy() for x in Vh]
    pack["L"] = [x.numpy() for x in L]
    return g, pack


def _hot(pack, g):
    from gpfy_b200.ops import HotPath
    m = [v.shape[0] for v in pack["V"]]
    return HotPath(int(g["meta_input_dim"]), m, num_outputs=pack["mu"].shape[1], kernel_order=pack["order"],
                   likelihood=pack["lik"], q_kind="tril" if "Lq" in pack else "fullcov",
                   weight_mode={0: "scalar", 1: "ard", 2: "projection"}[np.ndim(pack["w"])],
                   projection_dim=pack["w"].shape[1] if np.ndim(pack["w"]) == 2 else 0)


def _cat(pack):
    vhat = np.concatenate(pack["V"], 0)
    lcat = np.concatenate([l.ravel() for l in pack["L"]])
    return vhat, lcat


@pytest.mark.parametrize("name", ARD_CASES)
def test_features_match_reference_golden(name):
    g, pack = _resolved(name)
    hp = _hot(pack, g)
    vhat, lcat = _cat(pack)
    # polynomial_expansion on points already on the sphere (spherical_harmonics.py:290-320)
    Phi, _ = hp.features(g["o_x_sphere"], vhat, lcat, apply_to_sphere=False)
    assert rel_err(Phi.cpu().numpy(), g["o_Phi"]) < TOL
    # Kuf (spherical_harmonics.py:344-362) from raw X
    F = len(pack["V"])
    Kuf, r = hp.features(g["X"], vhat, lcat, w=pack["w"], b=pack["b"], degree_scale=np.full(F, np.sqrt(pack["v"])), mul_r=True)
    assert rel_err(Kuf.cpu().numpy(), g["o_Kuf"]) < TOL
    assert rel_err(r.cpu().numpy(), g["o_r"]) < TOL
    # projection A = sqrt(1/Kuu) * Kuf (spherical_harmonics.py:385-387)
    A, _ = hp.features(g["X"], vhat, lcat, w=pack["w"], b=pack["b"], degree_scale=np.sqrt(pack["eig"] * pack["v"]), mul_r=True)
    assert rel_err(A.cpu().numpy(), g["o_proj"]) < TOL
    # caller-provided output buffers are filled in place; wrong shapes are refused
    import torch
    from gpfy_b200 import _lib
    bufs = (torch.full_like(A, float("nan")), torch.full((A.shape[1],), float("nan"), dtype=torch.float64, device=A.device))
    A2, r2 = hp.features(g["X"], vhat, lcat, w=pack["w"], b=pack["b"], degree_scale=np.sqrt(pack["eig"] * pack["v"]), mul_r=True, out=bufs)
    assert A2.data_ptr() == bufs[0].data_ptr() and torch.equal(A2, A) and rel_err(r2.cpu().numpy(), g["o_r"]) < TOL
    with pytest.raises(_lib.GpfyError):
        hp.features(g["X"], vhat, lcat, w=pack["w"], b=pack["b"], out=(bufs[0][:, :-1], bufs[1]))


@pytest.mark.parametrize("name", [c for c in ARD_CASES if "proj" not in c])
def test_features_vjp_matches_oracle_autograd(name):
    """gpfy_sh_features_vjp: the feature matrix is differentiable on its own, as under jax.grad in the reference
    (polynomial_expansion / Kuf / project_fun, spherical_harmonics.py:290-320, 344-362, 385-387).  Cotangents of Vhat, the
    Cholesky factors, the per-degree scale, weight_variances and bias_variance against torch autograd on the oracle."""
    g, pack = _resolved(name)
    hp = _hot(pack, g)
    vhat, lcat = _cat(pack)
    F = len(pack["V"])
    rng = np.random.default_rng(3)
    M, N = hp.M, g["X"].shape[0]
    dPhi = rng.standard_normal((M, N))
    scale = np.sqrt(pack["eig"] * pack["v"])
    # raw X, projection A = scale_l r Phi
    got = hp.features_vjp(g["X"], vhat, lcat, dPhi, w=pack["w"], b=pack["b"], degree_scale=scale, mul_r=True)
    V = [O.T(v).requires_grad_(True) for v in pack["V"]]
    L = [O.T(l).requires_grad_(True) for l in pack["L"]]
    w, b, sc = O.T(pack["w"]).requires_grad_(True), O.T(pack["b"]).requires_grad_(True), O.T(scale).requires_grad_(True)
    Xs, r = O.to_sphere(w, b, O.T(g["X"]))
    Phi = O.polynomial_expansion(V, L, pack["alpha"], Xs)
    rows = torch.cat([sc[l].expand(v.shape[0]) for l, v in enumerate(pack["V"])])
    (torch.sum(O.T(dPhi) * Phi * r[None, :] * rows[:, None])).backward()
    assert rel_err(got["vhat"].cpu().numpy(), np.concatenate([v.grad.numpy() for v in V], 0)) < TOL
    assert rel_err(got["lcat"].cpu().numpy(), np.concatenate([np.tril(l.grad.numpy()).ravel() for l in L])) < TOL
    assert rel_err(got["scale"].cpu().numpy(), sc.grad.numpy()) < TOL
    assert rel_err(got["w"].cpu().numpy(), np.atleast_1d(w.grad.numpy()).ravel()) < TOL
    assert rel_err(got["b"].cpu().numpy().ravel(), np.atleast_1d(b.grad.numpy())) < TOL
    # points already on the sphere (what polynomial_expansion receives), no scale, no radius
    got2 = hp.features_vjp(g["o_x_sphere"], vhat, lcat, dPhi, apply_to_sphere=False)
    V2 = [O.T(v).requires_grad_(True) for v in pack["V"]]
    L2 = [O.T(l).requires_grad_(True) for l in pack["L"]]
    (torch.sum(O.T(dPhi) * O.polynomial_expansion(V2, L2, pack["alpha"], O.T(g["o_x_sphere"])))).backward()
    assert rel_err(got2["vhat"].cpu().numpy(), np.concatenate([v.grad.numpy() for v in V2], 0)) < TOL
    assert rel_err(got2["lcat"].cpu().numpy(), np.concatenate([np.tril(l.grad.numpy()).ravel() for l in L2])) < TOL


@pytest.mark.parametrize("name", ARD_CASES)
def test_predict_diag_matches_reference_golden(name):
    g, pack = _resolved(name)
    hp = _hot(pack, g)
    vhat, lcat = _cat(pack)
    sigma = pack.get("Lq", pack.get("Sigma"))
    fmu, fvar = hp.predict_diag(g["X"], pack["w"], pack["b"], pack["v"], pack["eig"], vhat, lcat, pack["mu"], sigma)
    assert rel_err(fmu.cpu().numpy(), g["o_fmu"]) < TOL
    assert rel_err(fvar.cpu().numpy(), g["o_fvar"]) < TOL


@pytest.mark.parametrize("name", ARD_CASES)
def test_elbo_value_and_gradients_match_oracle(name):
    g, pack = _resolved(name)
    hp = _hot(pack, g)
    vhat, lcat = _cat(pack)
    sigma = pack.get("Lq", pack.get("Sigma"))
    packed = hp.elbo_valgrad_packed(g["X"], g["Y"], pack["w"], pack["b"], pack["v"], pack["eig"], vhat, lcat, pack["mu"], sigma,
                                    pack.get("sigma2"))
    torch.cuda.synchronize()
    out = {k: v.cpu().numpy() for k, v in hp.unpack(packed).items()}
    # value: sum_n var_exp_n of the REFERENCE (golden)
    assert abs(out["data_term"] - np.sum(g["o_var_exp"])) <= TOL * abs(np.sum(g["o_var_exp"]))
    oval, og = O.elbo_and_grads(pack, g["X"], g["Y"], data_only=True)
    assert abs(out["data_term"] - oval) <= TOL * abs(oval)
    assert rel_err(out["mu"], og["mu"]) < TOL
    assert rel_err(out["sigma"], og["Lq"] if "Lq" in og else og["Sigma"]) < TOL
    assert rel_err(out["eig"], og["eig"]) < TOL
    assert rel_err(out["w"], np.atleast_1d(og["w"]).ravel()) < TOL
    assert rel_err(out["b"], og["b"]) < TOL
    assert rel_err(out["v"], og["v"]) < TOL
    if "sigma2" in og:
        assert rel_err(out["sigma2"], og["sigma2"]) < TOL
    gv = np.concatenate([og[f"V{i}"] for i in range(len(pack["V"]))], 0)
    assert rel_err(out["vhat"], gv) < TOL
    gl = np.concatenate([np.tril(og[f"L{i}"]).ravel() for i in range(len(pack["L"]))])
    assert rel_err(out["lcat"], gl) < TOL


def test_gegenbauer_matches_reference_golden():
    from gpfy_b200.ops import gegenbauer
    g = load("gegenbauer")
    for alpha in (0.5, 1.5, 3.0, 3.5, 10.0, 15.5):
        for n in (0, 1, 2, 5, 10, 12):
            c, dc = gegenbauer(n, alpha, g["x"], with_grad=True)
            assert rel_err(c.cpu().numpy(), g[f"C_{n}_{alpha}"]) < 1e-12
            ref = g[f"dC_{n}_{alpha}"]
            if n == 0:
                assert float(dc.abs().max()) == 0.0
            else:
                assert rel_err(dc.cpu().numpy(), ref) < 1e-12


@pytest.mark.parametrize("name", ["d8_F6_T12", "d5_ntk3_P2"])
def test_prior_kl_matches_reference_golden(name):
    g, pack = _resolved(name)
    hp = _hot(pack, g)
    kl, dmu, dsig = hp.prior_kl_valgrad(pack["mu"], pack["Lq"])
    assert abs(float(kl) - float(g["o_KL"])) <= TOL * abs(float(g["o_KL"]))
    mu_t = O.T(pack["mu"]).requires_grad_(True)
    L_t = O.T(pack["Lq"]).requires_grad_(True)
    O.prior_KL(mu_t, L_t, tril=True).backward()
    assert rel_err(dmu.cpu().numpy(), mu_t.grad.numpy()) < TOL
    # the reference differentiates the dense Sigma; entries above the diagonal get d/dL (1/2 L^2) = L = 0 there
    assert rel_err(dsig.cpu().numpy(), L_t.grad.numpy()) < TOL


def test_prior_kl_full_covariance_on_device_matches_reference_golden():
    """variational.py:375-399 on the device (gpfy_prior_kl_valgrad_ws): Cholesky logdet, trace and the symmetrised dKL/dSigma."""
    g, pack = _resolved("d4_fullcov_fixedV")
    hp = _hot(pack, g)
    kl, dmu, dsig = hp.prior_kl_valgrad(pack["mu"], pack["Sigma"])
    assert abs(float(kl) - float(g["o_KL"])) <= TOL * abs(float(g["o_KL"]))
    mu_t = O.T(pack["mu"]).requires_grad_(True)
    S_t = O.T(pack["Sigma"]).requires_grad_(True)
    O.prior_KL(mu_t, S_t, tril=False).backward()
    assert rel_err(dmu.cpu().numpy(), mu_t.grad.numpy()) < TOL
    sym = 0.5 * (S_t.grad.numpy() + np.swapaxes(S_t.grad.numpy(), -1, -2))
    assert rel_err(dsig.cpu().numpy(), sym) < TOL
    # a larger, worse-conditioned covariance: M = 280
    rng = np.random.default_rng(0)
    from gpfy_b200.ops import HotPath
    M = 280
    hp2 = HotPath(8, [1, 9] + [30] * 9, q_kind="fullcov")
    B = rng.standard_normal((M, M)) / np.sqrt(M)
    S = (B @ B.T + 0.05 * np.eye(M))[None]
    mu = rng.standard_normal((M, 1))
    kl2, dmu2, dsig2 = hp2.prior_kl_valgrad(mu, S)
    L = np.linalg.cholesky(S[0])
    ref = -0.5 * M - np.sum(np.log(np.diag(L))) + 0.5 * np.sum(mu ** 2) + 0.5 * np.trace(S[0])
    assert abs(float(kl2) - ref) <= TOL * abs(ref)
    assert rel_err(dsig2.cpu().numpy()[0], 0.5 * np.eye(M) - 0.5 * np.linalg.inv(S[0])) < 1e-10
