"""GPFY_PREC_F64_I8: the dense contraction C = W2 Z as an error-free split into signed radix-256 digits, multiplied on the int8
tensor cores (tcgen05 kind::i8, exact int32 TMEM accumulators) and recombined in float64 (csrc/gemm_i8.cuh).  It is a float64
path: held to the float64 contract of BASELINE.json north_star, relative 1e-10 against the reference goldens / the oracle, and
to 1e-11 against the DMMA float64 path on every block of the packed result (normwise per array, tests/golden_util.rel_err)."""
import numpy as np
import pytest
import torch

from tests.golden_util import CASES, load, rel_err
from tests.test_gpu_parity import _cat, _resolved

pytestmark = pytest.mark.gpu
TOL = 1e-10      # the float64 contract
TOL_DMMA = 1e-11  # against the DMMA path


def _hot(pack, g, precision):
    from gpfy_b200.ops import HotPath
    m = [v.shape[0] for v in pack["V"]]
    return HotPath(int(g["meta_input_dim"]), m, num_outputs=pack["mu"].shape[1], kernel_order=pack["order"], likelihood=pack["lik"],
                   q_kind="tril" if "Lq" in pack else "fullcov", weight_mode={0: "scalar", 1: "ard", 2: "projection"}[np.ndim(pack["w"])],
                   projection_dim=int(g["meta_projection_dim"]), precision=precision)


@pytest.mark.parametrize("name", CASES)
def test_i8_elbo_gradients_and_predictions_match_goldens_and_dmma(name):
    """Every golden case: shapes the int8 contraction is built for run it, the others (P = 2, projection weights) run the DMMA
    contraction under the same precision flag - either way the float64 contract holds."""
    g, pack = _resolved(name)
    vhat, lcat = _cat(pack)
    sigma = pack.get("Lq", pack.get("Sigma"))
    args = (g["X"], g["Y"], pack["w"], pack["b"], pack["v"], pack["eig"], vhat, lcat, pack["mu"], sigma, pack.get("sigma2"))
    h64, h8 = _hot(pack, g, "f64dmma"), _hot(pack, g, "f64i8")
    a = {k: v.cpu().numpy() for k, v in h64.unpack(h64.elbo_valgrad_packed(*args)).items()}
    b = {k: v.cpu().numpy() for k, v in h8.unpack(h8.elbo_valgrad_packed(*args)).items()}
    for k in a:
        if np.max(np.abs(a[k])) == 0.0:
            assert np.max(np.abs(b[k])) == 0.0, k
        else:
            assert rel_err(b[k], a[k]) < TOL_DMMA, (k, rel_err(b[k], a[k]))
    assert abs(b["data_term"] - np.sum(g["o_var_exp"])) <= TOL * abs(np.sum(g["o_var_exp"]))
    pa = (g["X"], pack["w"], pack["b"], pack["v"], pack["eig"], vhat, lcat, pack["mu"], sigma)
    fmu, fvar = h8.predict_diag(*pa)
    assert rel_err(fmu.cpu().numpy(), g["o_fmu"]) < TOL
    assert rel_err(fvar.cpu().numpy(), g["o_fvar"]) < TOL


@pytest.mark.parametrize("d,F,T,lik,N", [(8, 11, 30, "gaussian", 150_001), (8, 13, 30, "gaussian", 60_000), (32, 5, 64, "bernoulli", 60_000),
                                          (8, 11, 30, "bernoulli", 40_000), (2, 11, 12, "gaussian", 40_000), (8, 11, 20, "gaussian", 33_333), (8, 11, 64, "gaussian", 20_000),
                                          (8, 11, 100, "gaussian", 10_000)])
def test_i8_matches_dmma_at_the_bench_shapes(d, F, T, lik, N):
    """cfg3 / cfg4 / cfg5 shapes, S^2 and an M = 200 case (four out-blocks of 64 / 48 rows), tail tiles included."""
    from gpfy_b200.synthetic import headline_problem
    p64 = headline_problem(N, d=d, F=F, T=T, likelihood=lik, device="cuda", precision="f64dmma")
    p8 = headline_problem(N, d=d, F=F, T=T, likelihood=lik, device="cuda", precision="f64i8")
    a = p64["hot"].unpack(p64["hot"].elbo_valgrad_packed(*p64["args"]))
    b = p8["hot"].unpack(p8["hot"].elbo_valgrad_packed(*p8["args"]))
    for k in a:
        x, y = a[k].cpu().numpy(), b[k].cpu().numpy()
        if np.max(np.abs(x)) > 0:
            assert rel_err(y, x) < TOL_DMMA, (k, rel_err(y, x))
    pa = [p64["args"][i] for i in (0, 2, 3, 4, 5, 6, 7, 8, 9)]
    m64, v64 = p64["hot"].predict_diag(*pa)
    m8, v8 = p8["hot"].predict_diag(*pa)
    assert rel_err(m8.cpu().numpy(), m64.cpu().numpy()) < TOL_DMMA and rel_err(v8.cpu().numpy(), v64.cpu().numpy()) < TOL_DMMA


def test_i8_is_exactly_zero_at_the_prior():
    """S = I (L_q = I): W2 = 0, every digit is zero and C = 0 exactly, so fvar = K_diag with no rounding - the cancellation
    the reference's K_diag - sum proj^2 + sum (L^T proj)^2 reaches only approximately (model.py:129-131)."""
    from gpfy_b200.synthetic import headline_problem
    p = headline_problem(5000, device="cuda", precision="f64i8")
    args = list(p["args"])
    M = p["M"]
    args[9] = torch.eye(M, dtype=torch.float64, device="cuda").reshape(1, M, M)
    pa = [args[i] for i in (0, 2, 3, 4, 5, 6, 7, 8, 9)]
    p64 = headline_problem(5000, device="cuda", precision="f64dmma")
    _, v8 = p["hot"].predict_diag(*pa)
    _, v64 = p64["hot"].predict_diag(*pa)
    assert torch.equal(v8, v64)


def test_contraction_path_reports_the_arithmetic_actually_used():
    """gpfy_contraction_path: the int8 request is honoured where the kernel is built (P = 1, M_p <= 352, one of the fused backward
    kernels) and falls back to the DMMA contraction elsewhere - same float64 contract either way."""
    from gpfy_b200.ops import HotPath
    assert HotPath(8, [1, 9, 30, 30, 30]).contraction_path(100_000) == "f64i8"
    assert HotPath(8, [1, 9, 30, 30, 30], precision="f64dmma").contraction_path(100_000) == "f64"
    assert HotPath(8, [1, 9, 30, 30, 30], precision="f32").contraction_path(100_000) == "f32"
    assert HotPath(8, [1, 9, 30, 30, 30], num_outputs=2).contraction_path(100_000) == "f64"          # P > 1: legacy kernels
    assert HotPath(8, [1, 9] + [64] * 9).contraction_path(100_000) == "f64i8"                        # M = 586: out-blocks of 48 rows
    assert HotPath(8, [1, 9] + [100] * 9).contraction_path(100_000) == "f64"                         # M = 910 > 864: the W2 planes no longer fit
    assert HotPath(8, [1, 9, 30, 30, 30], weight_mode="projection", projection_dim=3).contraction_path(100_000) == "f64"
