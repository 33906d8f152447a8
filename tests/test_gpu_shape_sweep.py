"""Shape sweep of the fused step against the oracle: odd input dims, a single degree, one row per degree, many outputs,
both likelihoods, row counts around the 32-point tiles - every dispatch boundary of the kernels (ambient dim <= 16 or
not, M_p <= 288 or not, P = 1 or not, Gaussian fused in the backward kernel or not)."""
import numpy as np
import pytest
import torch

from oracle import gpfy_oracle as O
from tests.golden_util import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-10

SHAPES = [
    # d, F, T, P, lik, N
    (2, 1, 30, 1, "gaussian", 33),          # a single degree (the constant harmonic only)
    (2, 2, 1, 1, "gaussian", 64),           # one row per degree
    (3, 4, 2, 3, "bernoulli", 31),
    (4, 6, 7, 1, "bernoulli", 257),
    (5, 3, 100, 2, "gaussian", 1),          # one data point
    (7, 11, 30, 1, "gaussian", 95),
    (8, 11, 33, 1, "gaussian", 200),        # M_p = 320 > 288: general backward path for P = 1
    (11, 5, 16, 1, "gaussian", 129),        # ambient dim 12: the widest DIMP = 12 case
    (13, 4, 30, 1, "gaussian", 63),         # ambient dim 14 -> DIMP 16
    (15, 4, 12, 4, "gaussian", 100),        # ambient dim 16 exactly, P = 4
    (16, 4, 12, 1, "gaussian", 77),         # ambient dim 17 -> DIMP 20: first shape beyond the fused kernels
    (17, 3, 9, 1, "bernoulli", 40),
    (24, 3, 20, 2, "gaussian", 65),
    (31, 3, 5, 1, "gaussian", 32),          # ambient dim 32
    (32, 5, 30, 1, "bernoulli", 150),       # BASELINE config 5 at T = 30
    (8, 13, 12, 8, "gaussian", 50),         # P = 8 (the largest unrolled output count)
]


@pytest.mark.parametrize("d,F,T,P,lik,N", SHAPES)
def test_step_and_prediction_match_oracle(d, F, T, P, lik, N):
    from gpfy_b200.synthetic import headline_problem
    prob = headline_problem(N, d=d, F=F, T=T, P=P, likelihood=lik, device="cuda", seed=d * 100 + F)
    hp, args, host = prob["hot"], prob["args"], prob["host"]
    out = hp.unpack(hp.elbo_valgrad_packed(*args))
    pack = {"dim": host["dim"], "alpha": host["alpha"], "V": host["V"], "L": host["L"], "w": host["w"], "b": host["b"],
            "v": host["v"], "eig": host["eig"], "order": 1, "mu": host["mu"], "Lq": host["Lq"], "sigma2": host["sigma2"], "lik": lik}
    oval, og = O.elbo_and_grads(pack, args[0].cpu().numpy(), args[1].cpu().numpy(), data_only=True)
    assert abs(float(out["data_term"]) - oval) <= TOL * abs(oval)
    nF = len(host["V"])
    cmp = {"mu": og["mu"], "sigma": og["Lq"], "eig": og["eig"], "w": og["w"], "b": og["b"], "v": og["v"],
           "vhat": np.concatenate([og[f"V{i}"] for i in range(nF)], 0),
           "lcat": np.concatenate([np.tril(og[f"L{i}"]).ravel() for i in range(nF)])}
    if lik == "gaussian":
        cmp["sigma2"] = og["sigma2"]
    for k, ref in cmp.items():
        assert rel_err(out[k].cpu().numpy(), ref) < TOL, k
    fmu, fvar = hp.predict_diag(*[args[i] for i in (0, 2, 3, 4, 5, 6, 7, 8, 9)])
    ofmu, ofvar = O.predict_diag(O.pack_to_torch(pack), O.T(args[0].cpu().numpy()))
    assert rel_err(fmu.cpu().numpy(), ofmu.numpy()) < TOL and rel_err(fvar.cpu().numpy(), ofvar.numpy()) < TOL
    A, _ = hp.features(args[0], args[6], args[7], w=args[2], b=args[3], degree_scale=torch.sqrt(args[5] * args[4]), mul_r=True)
    assert rel_err(A.cpu().numpy(), O.projection(O.pack_to_torch(pack), O.T(args[0].cpu().numpy())).numpy()) < TOL
