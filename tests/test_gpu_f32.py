"""The f32 instantiation (GPFY_PREC_TF32X3: the dense contraction C = W2 Z as 3 x TF32 on tcgen05 with TMEM accumulators,
TMA tensor loads) against the float64 path: BASELINE.json north_star asks for relative 1e-5 in fp32.  The reference itself
is float64-only (utils.py:26, param.py:151), so the float64 path - pinned to the reference goldens elsewhere - is the oracle
here; tolerance 1e-5 normwise per array, as in tests/golden_util.rel_err."""
import numpy as np
import pytest
import torch

from tests.golden_util import CASES, load, rel_err
from tests.test_gpu_parity import _cat, _resolved

pytestmark = pytest.mark.gpu
TOL32 = 1e-5


def _hot(pack, g, precision):
    from gpfy_b200.ops import HotPath
    m = [v.shape[0] for v in pack["V"]]
    return HotPath(int(g["meta_input_dim"]), m, num_outputs=pack["mu"].shape[1], kernel_order=pack["order"], likelihood=pack["lik"],
                   q_kind="tril" if "Lq" in pack else "fullcov", weight_mode={0: "scalar", 1: "ard"}[np.ndim(pack["w"])],
                   precision=precision)


def _f32_cases():
    out = []
    for name in CASES:
        g = load(name)
        if int(g["meta_P"]) == 1 and int(g["meta_projection_dim"]) == 0:
            out.append(name)
    return out


@pytest.mark.parametrize("name", _f32_cases())
def test_f32_elbo_gradients_and_predictions_match_f64_on_goldens(name):
    g, pack = _resolved(name)
    vhat, lcat = _cat(pack)
    sigma = pack.get("Lq", pack.get("Sigma"))
    args = (g["X"], g["Y"], pack["w"], pack["b"], pack["v"], pack["eig"], vhat, lcat, pack["mu"], sigma, pack.get("sigma2"))
    h64, h32 = _hot(pack, g, "f64dmma"), _hot(pack, g, "f32")
    a = {k: v.cpu().numpy() for k, v in h64.unpack(h64.elbo_valgrad_packed(*args)).items()}
    b = {k: v.cpu().numpy() for k, v in h32.unpack(h32.elbo_valgrad_packed(*args)).items()}
    for k in a:
        if np.max(np.abs(a[k])) == 0.0:
            assert np.max(np.abs(b[k])) == 0.0, k
        else:
            assert rel_err(b[k], a[k]) < TOL32, (k, rel_err(b[k], a[k]))
    # the reference's own outputs (goldens), at the f32 tolerance
    assert abs(b["data_term"] - np.sum(g["o_var_exp"])) <= TOL32 * abs(np.sum(g["o_var_exp"]))
    pa = (g["X"], pack["w"], pack["b"], pack["v"], pack["eig"], vhat, lcat, pack["mu"], sigma)
    fmu, fvar = h32.predict_diag(*pa)
    assert rel_err(fmu.cpu().numpy(), g["o_fmu"]) < TOL32
    assert rel_err(fvar.cpu().numpy(), g["o_fvar"]) < TOL32


@pytest.mark.parametrize("d,F,T,lik,N", [(8, 11, 30, "gaussian", 150_001), (8, 13, 30, "gaussian", 60_000), (32, 5, 64, "bernoulli", 60_000),
                                          (8, 11, 30, "bernoulli", 40_000), (2, 11, 12, "gaussian", 40_000)])
def test_f32_matches_f64_at_the_bench_shapes(d, F, T, lik, N):
    """cfg3 / cfg4 / cfg5 shapes (BASELINE.json configs[2..4]) and S^2: every block of the packed result within 1e-5 of the
    float64 path, tail tiles included (N is not a multiple of the 128-point tcgen05 tile).  (The S^2 case is phase-truncated:
    with a COMPLETE degree of random points the synthetic Gram matrices are ill-conditioned and dSigma = E G E^T amplifies the
    fp32-level error of G by cond(L_l)^2 - 4.5e-4 measured - which no fp32 evaluation can avoid; the reference's own
    well-conditioned fundamental systems on S^2 are covered by the golden cases above.)"""
    from gpfy_b200.synthetic import headline_problem
    p64 = headline_problem(N, d=d, F=F, T=T, likelihood=lik, device="cuda", precision="f64dmma")
    p32 = headline_problem(N, d=d, F=F, T=T, likelihood=lik, device="cuda", precision="f32")
    a = p64["hot"].unpack(p64["hot"].elbo_valgrad_packed(*p64["args"]))
    b = p32["hot"].unpack(p32["hot"].elbo_valgrad_packed(*p32["args"]))
    for k in a:
        x, y = a[k].cpu().numpy(), b[k].cpu().numpy()
        if np.max(np.abs(x)) > 0:
            assert rel_err(y, x) < TOL32, (k, rel_err(y, x))
    pa = [p64["args"][i] for i in (0, 2, 3, 4, 5, 6, 7, 8, 9)]
    m64, v64 = p64["hot"].predict_diag(*pa)
    m32, v32 = p32["hot"].predict_diag(*pa)
    assert rel_err(m32.cpu().numpy(), m64.cpu().numpy()) < TOL32 and rel_err(v32.cpu().numpy(), v64.cpu().numpy()) < TOL32


def test_f32_refuses_shapes_it_is_not_built_for():
    from gpfy_b200 import _lib
    from gpfy_b200.synthetic import headline_problem
    p = headline_problem(1000, T=12, F=6, P=2, device="cuda", precision="f32")
    with pytest.raises(_lib.GpfyError, match="TF32X3"):
        p["hot"].elbo_valgrad_packed(*p["args"])
