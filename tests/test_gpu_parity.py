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
