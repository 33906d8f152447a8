"""The barrier-free kernels (csrc/zonal_fwd3.cuh: one warp = one 32-point tile through all row tiles; csrc/features3.cuh: one warp =
one (tile, degree group) unit) against the oracle and against the CTA-per-tile kernels they replace (zonal_fwd2_kernel,
features2_kernel, the legacy two-kernel feature path), selected with the library's development switches.
zonal_fwd3_kernel only runs when a launch has at least 2 * 8 * SMs tiles (every resident warp gets one), so these cases are sized
above that; the golden-vector tests (test_gpu_parity) already go through features3_kernel, which has no size condition."""
import numpy as np
import pytest
import torch

from oracle import gpfy_oracle as O
from tests.golden_util import elem_err, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _rows_for_fwd3():
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    return 2 * 8 * sms * 32 + 37        # ragged: the last tile has 5 points


@pytest.mark.parametrize("lik,P,T,F,d", [("gaussian", 1, 30, 11, 8), ("bernoulli", 1, 30, 11, 8), ("gaussian", 2, 12, 6, 8), ("gaussian", 1, 30, 9, 2)])
def test_zonal_fwd3_step_matches_oracle_and_zonal_fwd2(lik, P, T, F, d):
    """One chunk with enough tiles for the barrier-free forward, a q far from the prior: value and EVERY gradient against the oracle
    (1e-10 normwise, 1e-9 componentwise), and against the same step through zonal_fwd2_kernel (1e-13: only the order of the
    mu2 . z sum differs)."""
    from gpfy_b200 import ops
    from gpfy_b200.synthetic import headline_problem
    N = _rows_for_fwd3()
    prob = headline_problem(N, d=d, F=F, T=T, P=P, likelihood=lik, device="cuda", seed=13)
    hp, args, host = prob["hot"], list(prob["args"]), prob["host"]
    rng = np.random.default_rng(7)
    M = hp.M
    host["mu"] = rng.standard_normal((M, P))
    host["Lq"] = np.stack([np.tril(0.6 * np.eye(M) + 0.8 * rng.standard_normal((M, M)) / np.sqrt(M)) for _ in range(P)])
    args[8] = torch.as_tensor(host["mu"], device="cuda")
    args[9] = torch.as_tensor(host["Lq"], device="cuda")
    new = hp.elbo_valgrad_packed(*args).clone()
    try:
        ops.debug_option("no_fwd3", 1)
        old = hp.elbo_valgrad_packed(*args).clone()
    finally:
        ops.debug_option("no_fwd3", 0)
    a, b = hp.unpack(new), hp.unpack(old)
    for k in a:
        x, y = a[k].cpu().numpy(), b[k].cpu().numpy()
        if np.max(np.abs(y)) > 0:
            assert rel_err(x, y) < 1e-13, (k, rel_err(x, y))
    out = {k: v.cpu().numpy() for k, v in a.items()}
    pack = {"dim": host["dim"], "alpha": host["alpha"], "V": host["V"], "L": host["L"], "w": host["w"], "b": host["b"],
            "v": host["v"], "eig": host["eig"], "order": 1, "mu": host["mu"], "Lq": host["Lq"], "sigma2": host["sigma2"], "lik": lik}
    oval, og = O.elbo_and_grads(pack, args[0].cpu().numpy(), args[1].cpu().numpy(), data_only=True)
    assert abs(out["data_term"] - oval) <= TOL * abs(oval)
    nF = len(host["V"])
    ref = {"mu": og["mu"], "sigma": og["Lq"], "eig": og["eig"], "w": og["w"], "b": og["b"], "v": og["v"],
           "vhat": np.concatenate([og[f"V{i}"] for i in range(nF)], 0),
           "lcat": np.concatenate([np.tril(og[f"L{i}"]).ravel() for i in range(nF)])}
    if lik == "gaussian":
        ref["sigma2"] = og["sigma2"]
    for k, r in ref.items():
        assert rel_err(out[k].reshape(np.shape(r)), r) < TOL, k
        assert elem_err(out[k].reshape(np.shape(r)), r) < 1e-9, k


def test_zonal_fwd3_predictions_match_zonal_fwd2():
    from gpfy_b200 import ops
    from gpfy_b200.synthetic import headline_problem
    prob = headline_problem(_rows_for_fwd3(), device="cuda", seed=3)
    hp, args = prob["hot"], prob["args"]
    pa = [args[i] for i in (0, 2, 3, 4, 5, 6, 7, 8, 9)]
    m3, v3 = (t.clone() for t in hp.predict_diag(*pa))
    try:
        ops.debug_option("no_fwd3", 1)
        m2, v2 = (t.clone() for t in hp.predict_diag(*pa))
    finally:
        ops.debug_option("no_fwd3", 0)
    assert rel_err(m3.cpu().numpy(), m2.cpu().numpy()) < 1e-13 and rel_err(v3.cpu().numpy(), v2.cpu().numpy()) < 1e-13


@pytest.mark.parametrize("d,F,T,N", [(8, 11, 30, 20_011), (8, 13, 30, 20_011), (2, 9, 30, 5_003), (16, 6, 30, 4_097), (8, 11, 20, 33), (31, 4, 32, 1_001)])
@pytest.mark.parametrize("mul_r", [True, False])
def test_features3_equals_the_cta_per_tile_feature_kernels(d, F, T, N, mul_r):
    """Phi / Kuf through features3_kernel (triangles of Linv in shared memory where they fit, 13 degrees: through L1) and through the
    kernels it replaces: the same DMMA products in the same order, so bit-identical to features2_kernel where that one takes the shape,
    1e-13 against the legacy two-kernel path otherwise; radii identical."""
    from gpfy_b200 import ops
    from gpfy_b200.synthetic import headline_problem
    prob = headline_problem(N, d=d, F=F, T=T, device="cuda", seed=5)
    hp, args = prob["hot"], prob["args"]
    ds = torch.sqrt(args[5] * args[4])
    f = lambda: hp.features(args[0], args[6], args[7], w=args[2], b=args[3], degree_scale=ds if mul_r else None, mul_r=mul_r)
    new, rnew = (t.clone() if t is not None else None for t in f())
    try:
        ops.debug_option("no_features3", 1)
        old, rold = (t.clone() if t is not None else None for t in f())
    finally:
        ops.debug_option("no_features3", 0)
    assert torch.isfinite(new).all()
    assert rel_err(new.cpu().numpy(), old.cpu().numpy()) < 1e-13
    if rnew is not None:
        assert torch.equal(rnew, rold)
